"""Shared helpers for the parity tests (tests only)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PASS_CASES = ["pass_bias_dense", "pass_nobias_sparse", "pass_messy"]


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: g[k] for k in g.files}


def golden_bias_dict(g):
    from oracle import fithic_oracle as fo
    if not bool(g["has_bias"]):
        return None
    return fo.read_bias_arrays(g["bias_chrom"], g["bias_mid"], g["bias_val"])[0]


def log10_close(a, b, tol):
    """|log10 a - log10 b| <= tol where both > 0; exact match where either is 0 or 1."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    pos = (a > 0) & (b > 0)
    bad = np.zeros(a.shape, bool)
    with np.errstate(divide="ignore", invalid="ignore"):
        bad[pos] = np.abs(np.log10(a[pos]) - np.log10(b[pos])) > tol
    bad[~pos] = a[~pos] != b[~pos]
    return not bad.any(), int(bad.sum())
