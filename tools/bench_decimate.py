"""Throughput of K7 (bbk_decimate) on device-resident columns.

    python tools/bench_decimate.py [rows]        (needs a GPU)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from blueberry_b200 import _lib


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
    lib = _lib.load()
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(5)
    nb, K = 49851, 2000
    i = torch.randint(0, nb, (n,), device=dev, generator=gen)
    i, _ = torch.sort(i)                                         # row-major like a significances file
    d = torch.randint(0, K, (n,), device=dev, generator=gen)
    rows = torch.empty((n, 5), dtype=torch.float64, device=dev)
    rows[:, 0] = (i * 1000 + 500).double()
    rows[:, 1] = ((i + d) * 1000 + 500).double()
    rows[:, 2] = torch.randint(0, 40, (n,), device=dev, generator=gen).double()
    rows[:, 3] = torch.rand(n, device=dev, generator=gen, dtype=torch.float64) ** 2
    rows[:, 4] = torch.rand(n, device=dev, generator=gen, dtype=torch.float64)
    out = torch.empty((n, 5), dtype=torch.float64, device=dev)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = torch.empty(int(lib.bbk_decimate_workspace_bytes(n)), dtype=torch.uint8, device=dev)

    def run():
        _lib.check(lib.bbk_decimate(_lib.ptr(rows), n, 5000, _lib.ptr(out), _lib.ptr(n_out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "bbk_decimate")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    g = int(n_out.item())
    print("bbk_decimate: %d rows -> %d groups, %.2f ms, %.2f G rows/s, %.0f GB/s of the 40 B/row read + 40 B/group written"
          % (n, g, ms, n / ms / 1e6, (40.0 * n + 40.0 * g) / ms / 1e6))


if __name__ == "__main__":
    main()
