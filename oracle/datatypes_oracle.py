"""CPU restatement of the FithicContactMap methods next to the pass (blueberry/datatypes.pyx:274-350).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this.  Pinned against the reference's own method bodies executed in the build container
(oracle/run_reference.py: run_reference_decimate; tests/test_oracle_vs_reference.py) and frozen in
tests/golden/decimate.npz.

The reference is Python-2 code: `/` on the integer midpoints is floor division there (datatypes.pyx:331), which this
restatement writes as `//` - the same kind of edit as the closed list for fithic.py in oracle/ref_loader.py.
"""
import numpy as np

Q_LOWER_BOUND = 0.01        # utils.py:23


def decimate(map5, resolution=5000):
    """datatypes.pyx:317-339.  map5: (n, 5) float64 rows (mid1, mid2, contactCount, p, q).  Returns the decimated (g, 5)
    map, groups in the order of their first row (the iteration order of a Python-3 dict; Python 2's was arbitrary)."""
    m = np.array(map5, dtype=np.float64, copy=True)
    if len(m) == 0:
        return m.reshape(0, 5)
    m[:, :2] = (m[:, :2].astype('int') + resolution) // resolution * resolution - resolution // 2      # :331
    contact_values = {}
    for mid1, mid2, contactCount, p, q in m:                                                          # :334-337
        key = mid1, mid2
        contact0, p0, q0 = contact_values.get(key, (0, 1, 1))
        contact_values[key] = contactCount + contact0, p * p0, min(q, q0)
    return np.array([[mid1, mid2, contactCount, p, q] for (mid1, mid2), (contactCount, p, q) in contact_values.items()])


def contacts(map5):
    """datatypes.pyx:341-350: the midpoint pairs with q <= Q_LOWER_BOUND."""
    m = np.asarray(map5)
    return m[m[:, 4] <= Q_LOWER_BOUND, :2]


def regions(map5):
    """datatypes.pyx:315,339: every midpoint that takes part in a contact."""
    m = np.asarray(map5)
    return np.union1d(m[:, 0], m[:, 1])


def contact_map_dense(pos1, pos2, count, n_bins, resolution):
    """datatypes.pyx:99-120: the dense (n_bins+1)^2 matrix the reference builds from the RAWobserved rows
    (data = nan_to_num(data); j = int(pos1 / resolution); both triangles; a repeated cell keeps the last row)."""
    d = n_bins + 1
    data = np.nan_to_num(np.stack([np.asarray(pos1, dtype=np.float64), np.asarray(pos2, dtype=np.float64),
                                   np.asarray(count, dtype=np.float64)], axis=1))
    matrix = np.zeros((d, d), dtype=np.float64)
    for a, b, c in data:
        j, k = int(a / resolution), int(b / resolution)
        matrix[j, k] = c
        matrix[k, j] = c
    regions = np.union1d(data[:, 0], data[:, 1])
    return matrix, regions


def normalize_dense(matrix, kr_norm, kr_expected, n_bins):
    """datatypes.pyx:143-171: KR balancing and observed/expected in place, then nan_to_num.  The reference's division is
    Cython's checked C division: a divisor of exactly 0.0 raises ZeroDivisionError (NaN divisors just give NaN -> 0)."""
    m = np.array(matrix, dtype=np.float64, copy=True)
    with np.errstate(invalid="ignore"):
        for i in range(n_bins):
            for j in range(n_bins - i):
                den = kr_norm[j] * kr_norm[j + i] * kr_expected[i]
                if den == 0.0:
                    raise ZeroDivisionError("float division")
                m[j, j + i] /= den
                m[j + i, j] = m[j, j + i]
    return np.nan_to_num(m)


HIGH_FITHIC_CUTOFF = 10000000   # utils.py:25
LOW_FITHIC_CUTOFF = 25000       # utils.py:26


def extract_contacts(map5, chromosome, alpha=None):
    """utils.extract_contacts (utils.py:69-84) on an in-memory table: p <= alpha (:72-73), columns shifted right with the
    chromosome in front (:76-77), band filter on mid2 - mid1 (:80-83).  Returns (k, 5) rows (chromosome, mid1, mid2, count, p).
    Pinned by executing the reference's own function (oracle/run_reference.py: run_reference_extract_contacts)."""
    contact = np.array(map5, dtype=np.float64, copy=True).reshape(-1, 5)
    if alpha is not None:
        contact = contact[contact[:, 3] <= alpha]
    contact[:, 1:] = contact[:, :-1].copy()
    contact[:, 0] = chromosome
    distances = contact[:, 2] - contact[:, 1]
    return contact[(distances <= HIGH_FITHIC_CUTOFF) & (distances >= LOW_FITHIC_CUTOFF)]


def count_band_regions(regions):
    """blueberry.pyx:77-91: pairs (i, j < i) with LOW <= regions[i] - regions[j] <= HIGH."""
    r = np.asarray(regions, dtype=np.float64)
    t = 0
    for i in range(len(r)):
        d = r[i] - r[:i]
        t += int(((d >= LOW_FITHIC_CUTOFF) & (d <= HIGH_FITHIC_CUTOFF)).sum())
    return t


def benjamini_hochberg_sorted(p_sorted, n):
    """blueberry.pyx:40-75 on ALREADY sorted p: q[i] = max(q[i-1], min(p[i] * n / (i + 1), 1)) with Python's min / max."""
    q = np.zeros(len(p_sorted))
    prev = 0.0
    for i, p in enumerate(np.asarray(p_sorted, dtype=np.float64)):
        bh = p * n / (i + 1)
        bh = min(bh, 1)
        bh = max(bh, prev)
        prev = bh
        q[i] = bh
    return q


def genome_qvalues(maps, alpha=None):
    """The composition of SURVEY.md 3.2: extract every chromosome, n = sum of count_band_regions(union1d(mid1, mid2)),
    sort the p-values of all chromosomes, benjamini_hochberg(sorted p, n), q back in concatenation order.
    maps: {chromosome: (n, 5) table}.  Returns (contacts (m, 5), q (m,), n)."""
    parts, n = [], 0
    for chrom, m in maps.items():
        m = np.asarray(m, dtype=np.float64)
        n += count_band_regions(regions(m))
        parts.append(extract_contacts(m, chrom, alpha))
    contacts = np.concatenate(parts) if parts else np.zeros((0, 5))
    order = np.argsort(contacts[:, 4], kind="stable")
    q = np.empty(len(order))
    q[order] = benjamini_hochberg_sorted(contacts[order, 4], n)
    return contacts, q, n

