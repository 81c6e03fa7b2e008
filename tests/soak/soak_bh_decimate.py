"""One-off soak (needs a GPU): BH q-values / ranks at sizes around the tile and chunk edges of the cooperative rank kernel,
and decimate on random group structures, against the oracle."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

spec = importlib.util.spec_from_file_location("tp", os.path.join(ROOT, "tests", "test_gpu_parity.py"))
tp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tp)
from oracle import datatypes_oracle as do, fithic_oracle as fo
from blueberry_b200.datatypes import FithicContactMap

bad = 0
sizes = [1, 2, 31, 32, 33, 4095, 4096, 4097, 8191, 8192, 8193, 148 * 4096 - 1, 148 * 4096, 148 * 4096 + 1, 2 * 148 * 4096 + 5, 1_000_003]
for i, m in enumerate(sizes):
    rng = np.random.default_rng(700 + i)
    p = rng.random(m) ** 3
    p[rng.random(m) < 0.3] = 1.0
    if m > 100:
        p[rng.integers(0, m, m // 10)] = p[rng.integers(0, m, m // 10)]
        p[rng.integers(0, m, m // 30)] = np.nan
    for n_tests in (m, max(1, m // 50), 40 * m):
        q, rk = tp._bh_device(torch, p, n_tests, want_rank=True)
        valid = ~np.isnan(p)
        ref = fo.benjamini_hochberg_correction(p[valid], n_tests)
        srt = np.sort(p[valid])
        ok = np.isnan(q[~valid]).all() and np.array_equal(q[valid], ref) and np.array_equal(rk[valid], 1 + np.searchsorted(srt, p[valid], side="left"))
        q2, _ = tp._bh_device(torch, p, n_tests)
        ok = ok and np.array_equal(q2[valid], ref)
        if not ok:
            bad += 1
            print("BH FAILED m", m, "n_tests", n_tests)
print("BH sizes done, failures so far", bad)
for i in range(40):
    rng = np.random.default_rng(900 + i)
    n = int(rng.choice([1, 2, 33, 4096, 4097, 50_000, 148 * 4096 + 3, 300_000]))
    span = int(rng.choice([1, 3, 50, 2000, 40000]))
    m1 = rng.integers(0, span, n) * 1000 + 500
    m2 = m1 + rng.integers(0, max(span // 5, 1), n) * 1000
    mp = np.stack([m1, m2, rng.integers(0, 40, n), rng.random(n) ** 2, np.minimum(rng.random(n) * 2, 1.0)], axis=1).astype(np.float64)
    if rng.random() < 0.5:
        mp = mp[np.lexsort((mp[:, 1], mp[:, 0]))]
    r = int(rng.choice([2000, 5000, 25000]))
    cm = FithicContactMap.from_arrays(mp)
    cm.decimate(r)
    if not np.array_equal(cm.map, do.decimate(mp, r)):
        bad += 1
        print("decimate FAILED n", n, "span", span, "r", r)
print("total failures", bad)
