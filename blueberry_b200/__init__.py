"""blueberry_b200 - the Fit-Hi-C significance pass of jmschrei/blueberry on B200 (sm_100a).

Same entry points as the reference for this path (blueberry/fithic.py, and benjamini_hochberg /
count_band_regions of blueberry/blueberry.pyx); the work is done by hand-written CUDA kernels behind
the C ABI of include/bbk.h.  Importing the package does not need a GPU; calling a kernel entry point
without libbbk.so or without a CUDA device raises (there is no CPU fallback).
"""
from .utils import HIGH_FITHIC_CUTOFF, LOW_FITHIC_CUTOFF, Q_LOWER_BOUND, Q_UPPER_BOUND  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # lazy: torch is only imported when a kernel entry point is touched
    if name in ("FitHiC", "fithic", "benjamini_hochberg_correction", "generate_FragPairs", "read_interactions",
                "calculate_probabilities", "fit_spline", "read_bias_file", "in_range_check"):
        import importlib
        return getattr(importlib.import_module(__name__ + ".fithic"), name)
    if name in ("benjamini_hochberg", "count_band_regions"):
        import importlib
        return getattr(importlib.import_module(__name__ + ".blueberry"), name)
    raise AttributeError(name)
