"""Drop-in for the two Cython helpers of blueberry/blueberry.pyx that sit on the significance path.

    benjamini_hochberg(p_values, n) -> ndarray[float64]      blueberry.pyx:40-75  (input ALREADY sorted)
    count_band_regions(regions_ndarray) -> int               blueberry.pyx:77-91

Both run on the device through libbbk.so; there is no CPU fallback.
"""
import numpy as np
import torch

from . import _lib
from .utils import HIGH_FITHIC_CUTOFF, LOW_FITHIC_CUTOFF


def _device():
    if not torch.cuda.is_available():
        raise _lib.BbkError("blueberry_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def benjamini_hochberg(p_values, n):
    """Run the Benjamini-Hochberg procedure on a vector of -sorted- p-values (blueberry.pyx:40-75).

    q[i] = max(q[i-1], min(p[i] * n / (i+1), 1)): a forward running max over the given order; the
    input is taken as sorted and is not re-sorted, exactly like the reference.
    """
    dev = _device()
    lib = _lib.load()
    p = np.asarray(p_values).astype("float64")                  # blueberry.pyx:60
    m = int(p.shape[0])
    if m == 0:
        return np.zeros_like(p)
    dp = torch.empty((m + 1) & ~1, dtype=torch.float64, device=dev)[:m].copy_(torch.from_numpy(np.ascontiguousarray(p)))
    dq = torch.empty((m + 1) & ~1, dtype=torch.float64, device=dev)[:m]
    ws = torch.empty(int(lib.bbk_bh_workspace_bytes(m)), dtype=torch.uint8, device=dev)
    _lib.check(lib.bbk_bh_qvalues(_lib.ptr(dp), m, int(n), _lib.BH_POSITIONAL, None, _lib.ptr(dq), None,
                                  _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "bbk_bh_qvalues")
    return dq.cpu().numpy()


def count_band_regions(regions_ndarray, low=LOW_FITHIC_CUTOFF, high=HIGH_FITHIC_CUTOFF):
    """Number of region pairs (i, j<i) with low <= regions[i] - regions[j] <= high (blueberry.pyx:77-91).

    The reference reads the buffer as C doubles (blueberry.pyx:80), so the array must be float64.
    """
    dev = _device()
    lib = _lib.load()
    r = np.ascontiguousarray(regions_ndarray)
    if r.dtype != np.float64:
        raise TypeError("count_band_regions reads the buffer as C doubles (blueberry.pyx:80): pass a float64 array")
    n = int(r.shape[0])
    dr = torch.from_numpy(r).to(dev)
    res = torch.zeros(2, dtype=torch.int64, device=dev)
    _lib.check(lib.bbk_count_band(_lib.ptr(dr), n, float(int(low)), float(int(high)), _lib.ptr(res), _lib.stream_ptr()),
               "bbk_count_band")
    return int(res[0].item())
