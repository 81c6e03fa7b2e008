"""The fit stage's arithmetic (blueberry_b200/csrc/fit_stage.h, fit_coop.h), compiled for the HOST by
tests/host_harness, against scipy / sklearn / the oracle.  The product runs the very same source as an
sm_100a kernel; these tests pin the algorithm where no GPU is needed."""
import ctypes
import os
import warnings

import numpy as np
import pytest

from helpers import PASS_CASES, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DP = ctypes.POINTER(ctypes.c_double)


def _p(a, t=ctypes.c_double):
    return a.ctypes.data_as(ctypes.POINTER(t))


@pytest.fixture(scope="module")
def harness():
    import __graft_entry__
    __graft_entry__.build()
    return ctypes.CDLL(os.path.join(ROOT, "tests", "host_harness", "libfit_harness.so"))


def _fit(fn, x, y, s):
    m = len(x)
    n = ctypes.c_int()
    fp = ctypes.c_double()
    t = np.zeros(m + 4)
    c = np.zeros(m + 4)
    ier = fn(_p(x), _p(y), m, ctypes.c_double(s), ctypes.byref(n), ctypes.byref(fp), _p(t), _p(c))
    return ier, n.value, t[:n.value].copy(), c[:n.value].copy(), fp.value


def _datasets(n_trials, seed):
    rng = np.random.default_rng(seed)
    for trial in range(n_trials):
        m = int(rng.integers(5, 170))
        x = np.cumsum(rng.random(m) * 1e5 + 1)
        k = trial % 4
        if k == 0:
            y = 1e-3 * (x / 5000 + 1) ** -1.08 * np.exp(rng.normal(0, 0.05, m))
        elif k == 1:
            y = 1e-4 * (1.5 + np.sin(x / 3e5 * rng.random())) + rng.random(m) * 3e-5
        elif k == 2:
            y = 1e-3 * (x / 5000 + 1) ** -1.0 * np.exp(rng.normal(0, 0.3, m))
        else:
            y = 1e-5 * (x / 5000 + 1) ** -0.5 * (1 + 0.01 * rng.normal(0, 1, m))
        if trial % 29 == 0:
            y[rng.integers(0, m)] = 0.0                 # s = min(y)**2 = 0 -> interpolating spline
        s = float(min(y) ** 2)
        if trial % 5 == 1:
            s *= 1e-2
        yield x, y, s


def test_smoothing_spline_is_bit_identical_to_scipy(harness):
    from scipy.interpolate import UnivariateSpline
    iers = {}
    for x, y, s in _datasets(160, 1):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            u = UnivariateSpline(x, y, s=s)
        d = u._data
        n_s, t_s, c_s, ier_s = d[7], d[8][:d[7]], d[9][:d[7]], d[13]
        for fn in (harness.th_univariate_spline, harness.th_coop_univariate_spline):
            ier, n, t, c, fp = _fit(fn, x, y, s)
            assert (ier, n) == (ier_s, n_s)
            assert np.array_equal(t, t_s)
            assert np.array_equal(c[:n - 4], c_s[:n - 4])
        iers[ier_s] = iers.get(ier_s, 0) + 1
    assert iers.get(0, 0) > 50 and len(iers) >= 3       # converged, polynomial and interpolating cases all hit


def test_spline_evaluation_matches_scipy(harness):
    from scipy.interpolate import UnivariateSpline
    for x, y, s in _datasets(20, 2):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            u = UnivariateSpline(x, y, s=s)
        ier, n, t, c, fp = _fit(harness.th_coop_univariate_spline, x, y, s)
        xs = np.ascontiguousarray(np.linspace(x[0], x[-1], 777))
        out = np.zeros(len(xs))
        harness.th_spline_eval(_p(t), n, _p(np.ascontiguousarray(c)), _p(xs), len(xs), _p(out))
        assert np.array_equal(out, u(xs))


def test_antitonic_regression_is_bit_identical_to_sklearn(harness):
    from sklearn.isotonic import IsotonicRegression
    rng = np.random.default_rng(3)
    for trial in range(60):
        L = int(rng.integers(1, 2500))
        X = (np.arange(L) + 3) * 5000
        v = 1e-3 * (np.arange(L) + 1.0) ** -1.08 * (1 + 0.2 * rng.normal(0, 1, L) * (rng.random(L) < 0.3))
        if trial % 5 == 0:
            v[rng.integers(0, L, L // 3 + 1)] = v[0]
        ref = IsotonicRegression(increasing=False).fit_transform([int(a) for a in X], v)
        out = np.zeros(L)
        harness.th_antitonic(_p(np.ascontiguousarray(v)), L, _p(out))
        assert np.array_equal(out, ref)


def test_segmented_antitonic_regression_is_bit_identical_to_sklearn(harness):
    """The kernel's variant (pieces between safe cuts, run by 128 workers): smooth mostly-monotone curves, noisy ones,
    ties, plateaus, near-ties one ulp apart (no cut may be made there), increasing input (one big pool), NaN-free edge sizes."""
    from sklearn.isotonic import IsotonicRegression
    rng = np.random.default_rng(4)
    cases = []
    for trial in range(80):
        L = int(rng.integers(1, 3000))
        base = 1e-3 * (np.arange(L) + 1.0) ** -1.08
        kind = trial % 8
        if kind == 0:
            v = base                                                             # strictly decreasing: all singletons
        elif kind == 1:
            v = base * (1 + 0.2 * rng.normal(0, 1, L) * (rng.random(L) < 0.3))   # scattered wiggles
        elif kind == 2:
            v = base * (1 + 0.02 * np.sin(np.arange(L) / 7.0))                   # smooth oscillation
        elif kind == 3:
            v = base.copy(); v[rng.integers(0, L, L // 3 + 1)] = v[0]            # many exact ties
        elif kind == 4:
            v = np.sort(rng.random(L))                                           # increasing: one pool
        elif kind == 5:
            v = np.repeat(base[:max(L // 5, 1)], 5)[:L]                          # plateaus (ties pool)
            L = len(v)
        elif kind == 6:
            v = base.copy()                                                      # neighbours one ulp apart, both ways
            k = rng.integers(0, L, L // 4 + 1)
            v[k] = np.nextafter(v[np.maximum(k - 1, 0)], rng.choice([0.0, 1.0], len(k)))
        else:
            v = rng.normal(0, 1, L)                                              # noise, negative values
        cases.append(np.ascontiguousarray(v, dtype=np.float64))
    for v in cases:
        L = len(v)
        ref = IsotonicRegression(increasing=False).fit_transform(np.arange(L), v)
        for nt in (128, 7, 1):
            out = np.full(L, -1.0)
            harness.th_antitonic_segmented(_p(v), L, nt, _p(out))
            assert np.array_equal(out, ref), (L, nt)


@pytest.mark.parametrize("name", PASS_CASES)
def test_equal_occupancy_binning_is_bit_identical_to_the_reference(name, harness):
    from oracle import fithic_oracle as fo
    g = load_golden(name)
    poss = np.ascontiguousarray(g["ref_possible"])
    obs = np.ascontiguousarray(g["ref_observed"])
    nk = len(poss)
    x = np.zeros(1024)
    y = np.zeros(1024)
    bok = np.zeros(nk, np.int32)
    n = ctypes.c_int()
    st = harness.th_equal_occupancy(_p(poss, ctypes.c_int64), _p(obs, ctypes.c_int64), nk, ctypes.c_int64(int(g["ref_S"])),
                                    int(g["n_bins"]), ctypes.c_int64(int(g["resolution"])), ctypes.c_int64(int(g["ref_min_dist"])),
                                    ctypes.c_int64(int(g["ref_max_dist"])), _p(x), _p(y), 1024, _p(bok, ctypes.c_int32), ctypes.byref(n))
    assert st == 0
    assert np.array_equal(x[:n.value], g["ref_x"]) and np.array_equal(y[:n.value], g["ref_y"])
    _, _, _, obk = fo.calculate_probabilities(poss, obs, int(g["ref_S"]), int(g["n_bins"]), int(g["resolution"]),
                                              int(g["ref_min_dist"]), int(g["ref_max_dist"]))
    assert np.array_equal(bok, obk)
