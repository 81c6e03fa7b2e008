"""K5 alone: bbk_bh_qvalues (unsorted p -> q) on m random p-values that all need a rank (q < 1), time against m.

    python tools/bench_bh.py [m ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from blueberry_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda", 0)
sizes = [int(float(a)) for a in sys.argv[1:]] or [10 ** 4, 10 ** 5, 1100000, 10 ** 7]
for m in sizes:
    g = torch.Generator(device=dev); g.manual_seed(m)
    p = (torch.rand(m, generator=g, device=dev, dtype=torch.float64) * 1e-4)
    q = torch.empty_like(p)
    ws = torch.empty(int(lib.bbk_bh_workspace_bytes(m)), dtype=torch.uint8, device=dev)
    n_tests = m * 3

    def run():
        _lib.check(lib.bbk_bh_qvalues(_lib.ptr(p), m, n_tests, _lib.BH_UNSORTED, None, _lib.ptr(q), None, _lib.ptr(ws), ws.numel(),
                                      _lib.stream_ptr()), "bbk_bh_qvalues")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    below = int((q < 1.0).sum())
    print("m = %9d   %.3f ms   %.1f M keys/s   (q < 1 on %d rows)" % (m, ms, m / ms / 1e3, below))
