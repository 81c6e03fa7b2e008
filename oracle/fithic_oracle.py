"""CPU restatement of blueberry's Fit-Hi-C pass on ARRAYS.  TEST INFRASTRUCTURE ONLY.

Every function follows one reference function and cites it (paths relative to
/root/reference/).  Records are arrays instead of gzip text lines; chromosomes
are small integer ids.  Integer work is bit-exact by construction; the three
third-party calls the reference makes (scipy.special.bdtrc,
scipy.interpolate.UnivariateSpline, sklearn.isotonic.IsotonicRegression) are made
here too, on the same values in the same order, so this oracle reproduces the
reference's floating point exactly (checked by tests/test_oracle_vs_reference.py
against the reference itself and by tests/golden/*).

Parity: pinned against the reference executed in the build container
(oracle/ref_loader.py); the reference holds no tests/golden vectors of its own.

The product (blueberry_b200/) never imports this module.
"""
import bisect

import numpy as np
import scipy.special as scsp
from scipy.interpolate import UnivariateSpline
from sklearn.isotonic import IsotonicRegression

DIST_SCALING = 10000.0            # fithic.py:45
HIGH_FITHIC_CUTOFF = 10000000     # utils.py:25
LOW_FITHIC_CUTOFF = 25000         # utils.py:26


def in_range_check(d, min_dist, max_dist):
    """fithic.py:445-449 (vectorised): min < d <= max, -1 = unbounded."""
    d = np.asarray(d)
    lo = np.ones(d.shape, bool) if min_dist == -1 else (d > min_dist) & (min_dist > -1)
    hi = np.ones(d.shape, bool) if max_dist == -1 else (d <= max_dist) & (max_dist > -1)
    return lo & hi


class FragStats(object):
    """What generate_FragPairs leaves behind (fithic.py:272-332)."""
    __slots__ = ("resolution", "max_possible_dist", "possible", "possible_inter_all",
                 "possible_intra_all", "possible_intra_in_range", "n_frags")


def generate_frag_pairs(frag_chrom, frag_mid, resolution, min_dist, max_dist):
    """fithic.py:272-332.  possible[k] is mainDic[k*R][0]."""
    frag_chrom = np.asarray(frag_chrom)
    frag_mid = np.asarray(frag_mid, dtype=np.int64)
    R = int(resolution)
    st = FragStats()
    st.resolution = R
    chroms = np.unique(frag_chrom)
    per = []
    n_frags = 0
    max_possible = 0                                   # module global starts at 0 (:42)
    for c in chroms:
        mids = np.unique(frag_mid[frag_chrom == c])    # allFragsDic[chr] is a dict keyed by mid (:289-291)
        n = int(mids.size)
        max_frag = int(mids.max()) - R // 2            # :298 (py2 integer division)
        per.append((n, max_frag))
        n_frags += n
        max_possible = max(max_possible, max_frag)     # :300
    nkeys = len(range(0, max_possible + 1, R))         # :302
    possible = np.zeros(nkeys, dtype=np.int64)
    inter = 0
    intra = 0
    for n, max_frag in per:
        cnt = len(range(0, max_frag + 1, R))           # :309
        possible[:cnt] += n - np.arange(cnt, dtype=np.int64)   # :310-311
        inter += n * (n_frags - n)                     # :313
        intra += (n * (n + 1)) // 2                    # :314
    inter //= 2                                        # :316
    keys = np.arange(nkeys, dtype=np.int64) * R
    st.max_possible_dist = max_possible
    st.possible = possible
    st.possible_inter_all = inter
    st.possible_intra_all = intra
    st.possible_intra_in_range = int(possible[in_range_check(keys, min_dist, max_dist)].sum())  # :320-322
    st.n_frags = n_frags
    return st


class ContactStats(object):
    """What read_interactions leaves behind (fithic.py:229-270)."""
    __slots__ = ("observed", "S", "intra_in_range_count", "intra_all_sum", "intra_all_count",
                 "inter_all_sum", "inter_all_count", "min_obs_dist", "max_obs_dist")


def read_interactions(nkeys, resolution, chr1, mid1, chr2, mid2, count, min_dist, max_dist):
    """fithic.py:229-270.  observed[k] is mainDic[k*R][1]; chr1/chr2 may be None (= same chromosome)."""
    R = int(resolution)
    mid1 = np.asarray(mid1, dtype=np.int64)
    mid2 = np.asarray(mid2, dtype=np.int64)
    count = np.asarray(count, dtype=np.int64)
    d = mid2 - mid1                                     # :247 (no abs, no chromosome check)
    st = ContactStats()
    if chr1 is None:
        inter = np.zeros(d.shape, bool)
    else:
        inter = np.asarray(chr1) != np.asarray(chr2)    # :249
    st.inter_all_sum = int(count[inter].sum())
    st.inter_all_count = int(inter.sum())
    st.intra_all_sum = int(count[~inter].sum())
    st.intra_all_count = int((~inter).sum())
    rng = in_range_check(d, min_dist, max_dist)         # :256-257
    dr, cr = d[rng], count[rng]
    st.min_obs_dist = int(min(500000000, dr.min())) if dr.size else 500000000   # :40, :258
    st.max_obs_dist = int(max(0, dr.max())) if dr.size else 0                   # :41, :259
    iskey = (dr >= 0) & (dr % R == 0) & (dr // R < nkeys)                       # :260 "distance in mainDic"
    observed = np.zeros(nkeys, dtype=np.int64)
    np.add.at(observed, dr[iskey] // R, cr[iskey])      # :261
    st.observed = observed
    st.S = int(cr.sum())                                # :262 (counted even when d is not a key)
    st.intra_in_range_count = int(rng.sum())            # :263
    return st


def calculate_probabilities(possible, observed, S, n_bins, resolution, min_dist, max_dist):
    """fithic.py:160-227, literally (Python ints and floats, same operation order).

    Returns (x, y, yerr, bin_of_key) where bin_of_key[k] is the index of the emitted bin
    that distance k*R went into, -1 if none (out of range or in the dropped trailing bin).
    """
    R = int(resolution)
    possible = [int(v) for v in possible]
    observed = [int(v) for v in observed]
    S = int(S)
    desired = S // n_bins                               # :167 (py2 int floor)
    x, y, yerr = [], [], []
    bin_of_key = np.full(len(possible), -1, dtype=np.int32)
    acc = 0
    n = 0
    total = 0
    members = []
    for k in range(len(possible)):                      # :182
        i = k * R
        total += observed[k]                            # :183
        if not bool(in_range_check(i, min_dist, max_dist)):   # :184
            continue
        full = False
        if observed[k] >= desired:                      # :188
            members.append(k); acc = 0; full = True
        elif acc + observed[k] >= desired:              # :194
            members.append(k); acc = 0; full = True
        else:                                           # :199
            members.append(k); acc += observed[k]
        if full:                                        # :203
            n_pairs, n_inter, avg = 0.0, 0.0, 0.0
            n += 1
            if n < n_bins:
                desired = 1.0 * (S - total) / (n_bins - n)       # :209
            for b in members:                           # :211-214
                n_pairs += possible[b]
                n_inter += observed[b]
                avg += 1.0 * possible[b] * ((b * R) / DIST_SCALING)
            mean_prob = (n_inter / n_pairs) / S         # :216 (ZeroDivisionError if n_pairs == 0)
            avg = DIST_SCALING * (avg / n_pairs)        # :217
            bin_of_key[members] = len(x)
            x.append(avg); y.append(mean_prob); yerr.append(0.0)
            acc = 0
            members = []
    return x, y, yerr, bin_of_key


def fit_spline_tables(x, y, nkeys, resolution):
    """fithic.py:340-374: the fit part of fit_spline.

    Returns (spline_x0_index, newSplineY ndarray, residual, splineY_raw ndarray, ius):
    splineX is the key range k0..k0+L-1 (times R).
    """
    R = int(resolution)
    s = min(y) ** 2                                     # :340
    ius = UnivariateSpline(x, y, s=s)                   # :343
    min_x, max_x = min(x), max(x)                       # :350
    keys = [k * R for k in range(nkeys)]
    splineX = [i for i in keys if min_x <= i <= max_x]  # :351-357
    splineY = ius(splineX)                              # :359
    ir = IsotonicRegression(increasing=False)           # :361
    new_y = ir.fit_transform(splineX, splineY)          # :362
    residual = sum([i * i for i in (y - ius(x))])       # :374
    k0 = splineX[0] // R
    return k0, np.asarray(new_y, dtype=np.float64), residual, np.asarray(splineY, dtype=np.float64), ius


def spline_index(d, k0, L, resolution, min_x, max_x):
    """fithic.py:429-430 per record, literally (bisect on the splineX list)."""
    R = int(resolution)
    splineX = [(k0 + j) * R for j in range(L)]
    out = np.empty(len(d), dtype=np.int64)
    for n, dist in enumerate(d):
        t = min(max(int(dist), min_x), max_x)
        out[n] = min(bisect.bisect_left(splineX, t), L - 1)
    return out


def spline_index_closed_form(d, k0, L, resolution):
    """Closed form of fithic.py:429-430: clamp(ceil((d - splineX[0]) / R), 0, L-1)."""
    R = int(resolution)
    d = np.asarray(d, dtype=np.int64)
    q = -((-(d - k0 * R)) // R)
    return np.clip(q, 0, L - 1)


def bias_lookup(bias, chrom, mid):
    """fithic.py:418-425: biasDic[chr][mid] with default 1.0.  bias = dict{chrom: dict{mid: value}}."""
    out = np.ones(len(mid), dtype=np.float64)
    if not bias:
        return out
    if chrom is None:
        chrom = np.zeros(len(mid), dtype=np.int64)
    for c in np.unique(chrom):
        sub = bias.get(int(c))
        if not sub:
            continue
        sel = np.nonzero(np.asarray(chrom) == c)[0]
        keys = np.fromiter(sub.keys(), dtype=np.int64, count=len(sub))
        vals = np.fromiter(sub.values(), dtype=np.float64, count=len(sub))
        order = np.argsort(keys)
        keys, vals = keys[order], vals[order]
        m = np.asarray(mid)[sel].astype(np.int64)
        pos = np.clip(np.searchsorted(keys, m), 0, len(keys) - 1)
        hit = keys[pos] == m
        out[sel[hit]] = vals[pos[hit]]
    return out


def read_bias_arrays(bias_chrom, bias_mid, bias_val):
    """fithic.py:136-158 on arrays: out-of-[0.5,2] -> -1 (:147-149), first occurrence wins (:153-154)."""
    biases = {}
    discarded = 0
    for c, m, b in zip(bias_chrom, bias_mid, bias_val):
        c, m, b = int(c), int(m), float(b)
        if b < 0.5 or b > 2:
            b = -1
            discarded += 1
        sub = biases.setdefault(c, {})
        if m not in sub:
            sub[m] = b
    return biases, discarded


def score_pairs(mid1, mid2, count, S, k0, new_spline_y, resolution, min_dist, max_dist,
                bias1=None, bias2=None):
    """fithic.py:413-435 vectorised.  Returns (p, scored, keep):

    scored = min_dist <= d <= max_dist (:427, inclusive on both sides - unlike the stats pass);
    keep = scored & (p <= 1) (:434, drops NaN); p is NaN where not scored.
    """
    mid1 = np.asarray(mid1, dtype=np.int64)
    mid2 = np.asarray(mid2, dtype=np.int64)
    count = np.asarray(count, dtype=np.int64)
    d = mid2 - mid1
    L = len(new_spline_y)
    scored = (d >= min_dist) & (d <= max_dist)
    idx = spline_index_closed_form(d, k0, L, resolution)
    b1 = np.ones(len(d)) if bias1 is None else np.asarray(bias1, dtype=np.float64)
    b2 = np.ones(len(d)) if bias2 is None else np.asarray(bias2, dtype=np.float64)
    prior = np.asarray(new_spline_y)[idx] * (b1 * b2)   # :431
    p = np.full(len(d), np.nan)
    with np.errstate(all="ignore"):
        # scalar call in the reference is bdtrc(int, int, float) -> the 'dld' loop (int n)
        p[scored] = scsp.bdtrc((count[scored] - 1).astype(np.float64), np.int64(S), prior[scored])   # :432
    keep = scored & (p <= 1)
    return p, scored, keep


def benjamini_hochberg_correction(p_values, num_total_tests):
    """fithic.py:466-487 vectorised: FORWARD running max of min(p*N/rank, 1), input order out."""
    p = np.asarray(p_values, dtype=np.float64)
    order = p.argsort(kind="stable")
    sp = p[order]
    bh = sp * num_total_tests / np.arange(1, len(sp) + 1)      # two roundings, as :474
    bh = np.minimum(bh, 1)
    bh = np.maximum.accumulate(bh) if len(bh) else bh
    q = np.empty(len(sp), dtype=np.float64)
    q[order] = bh
    return q


def benjamini_hochberg_sorted(p_sorted, n):
    """blueberry.pyx:40-75 vectorised (input already sorted)."""
    p = np.asarray(p_sorted).astype("float64")
    bh = np.minimum(p * n / np.arange(1, len(p) + 1), 1)
    return np.maximum.accumulate(bh) if len(bh) else bh


def count_band_regions(regions, low=LOW_FITHIC_CUTOFF, high=HIGH_FITHIC_CUTOFF):
    """blueberry.pyx:77-91: #{(i, j<i): low <= regions[i]-regions[j] <= high}; O(n^2) blocked."""
    r = np.asarray(regions, dtype=np.float64)
    t = 0
    for lo in range(0, len(r), 2048):
        blk = r[lo:lo + 2048]
        diff = blk[:, None] - r[None, :lo + len(blk)]
        ok = (diff >= low) & (diff <= high)
        ok &= np.arange(lo, lo + len(blk))[:, None] > np.arange(lo + len(blk))[None, :]
        t += int(ok.sum())
    return t


class PassResult(object):
    __slots__ = ("frag", "contacts", "x", "y", "bin_of_key", "k0", "spline_y", "spline_y_raw",
                 "residual", "p", "scored", "keep")


def fithic_arrays(frag_chrom, frag_mid, chr1, mid1, chr2, mid2, count, resolution,
                  n_bins=100, min_dist=0, max_dist=10000000, bias=None):
    """fithic.py:110-133 on arrays (single pass, as the reference runs)."""
    res = PassResult()
    res.frag = generate_frag_pairs(frag_chrom, frag_mid, resolution, min_dist, max_dist)
    nkeys = len(res.frag.possible)
    res.contacts = read_interactions(nkeys, resolution, chr1, mid1, chr2, mid2, count, min_dist, max_dist)
    x, y, _, res.bin_of_key = calculate_probabilities(res.frag.possible, res.contacts.observed,
                                                     res.contacts.S, n_bins, resolution, min_dist, max_dist)
    res.x, res.y = x, y
    res.k0, res.spline_y, res.residual, res.spline_y_raw, _ = fit_spline_tables(x, y, nkeys, resolution)
    b1 = b2 = None
    if bias:
        b1 = bias_lookup(bias, chr1, mid1)
        b2 = bias_lookup(bias, chr2, mid2)
    res.p, res.scored, res.keep = score_pairs(mid1, mid2, count, res.contacts.S, res.k0, res.spline_y,
                                              resolution, min_dist, max_dist, b1, b2)
    return res


def fithic_two_pass_arrays(frag_chrom, frag_mid, chr1, mid1, chr2, mid2, count, resolution,
                           n_bins=100, min_dist=0, max_dist=10000000, bias=None):
    """Second pass (BASELINE config 4).  The reference has no pass-2 code (n_passes is ignored,
    fithic.py:121-133); SURVEY.md section 8c composes it from the reference's own functions:
    outliers = rows of pass 1 with p <= 1 / possibleIntraInRangeCount (fithic.py:320-322); the statistics
    (read_interactions, calculate_probabilities, the spline) are recomputed on the records that are not
    outliers, and fit_spline then scores ALL records with the refitted S and spline.
    Returns (pass1 PassResult, pass2 PassResult, outlier mask, threshold)."""
    r1 = fithic_arrays(frag_chrom, frag_mid, chr1, mid1, chr2, mid2, count, resolution, n_bins, min_dist, max_dist, bias)
    thr = 1.0 / r1.frag.possible_intra_in_range
    with np.errstate(invalid="ignore"):
        outlier = r1.keep & (r1.p <= thr)
    sel = ~outlier
    sub = lambda a: None if a is None else np.asarray(a)[sel]
    r2 = PassResult()
    r2.frag = r1.frag
    nkeys = len(r1.frag.possible)
    r2.contacts = read_interactions(nkeys, resolution, sub(chr1), sub(mid1), sub(chr2), sub(mid2), sub(count), min_dist, max_dist)
    x, y, _, r2.bin_of_key = calculate_probabilities(r2.frag.possible, r2.contacts.observed, r2.contacts.S, n_bins,
                                                     resolution, min_dist, max_dist)
    r2.x, r2.y = x, y
    r2.k0, r2.spline_y, r2.residual, r2.spline_y_raw, _ = fit_spline_tables(x, y, nkeys, resolution)
    b1 = b2 = None
    if bias:
        b1 = bias_lookup(bias, chr1, mid1)
        b2 = bias_lookup(bias, chr2, mid2)
    r2.p, r2.scored, r2.keep = score_pairs(mid1, mid2, count, r2.contacts.S, r2.k0, r2.spline_y, resolution,
                                           min_dist, max_dist, b1, b2)
    return r1, r2, outlier, thr
