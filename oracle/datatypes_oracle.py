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
