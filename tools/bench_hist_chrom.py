"""K1 on a shard WITH chromosome columns (20 B/record): the 32-bit fast path (min_dist >= 0) against the general path
(min_dist = -1 selects it) on the same 1e8 records.  Needs a GPU."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blueberry_b200.engine import PassEngine, Shard  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    n, R, nkeys = 100_000_000, 5000, 2001
    g = torch.Generator(device=dev).manual_seed(1)
    b1 = torch.randint(0, 49000, (n,), device=dev, generator=g, dtype=torch.int32)
    dk = torch.randint(0, 2001, (n,), device=dev, generator=g, dtype=torch.int32)
    mid1 = b1 * R + 2500
    mid2 = mid1 + dk * R
    count = (torch.rand(n, device=dev, generator=g) < 0.3).to(torch.int32) * torch.randint(1, 9, (n,), device=dev, generator=g, dtype=torch.int32)
    chrom = torch.zeros(n, dtype=torch.int32, device=dev)
    for label, cols, min_dist in (("chr columns, fast path", (chrom, chrom), 0), ("chr columns, general path", (chrom, chrom), -1),
                                  ("compact, fast path", (None, None), 0), ("compact, general path", (None, None), -1)):
        sh = Shard(mid1, mid2, count, cols[0], cols[1])
        eng = PassEngine(R, 100, min_dist, 10_000_000, nkeys, device=dev)
        for _ in range(3):
            eng.hist([sh])
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            eng.hist([sh])
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        bpr = 20 if cols[0] is not None else 12
        print("%-28s %.3f ms  %.0f GB/s  (S = %d)" % (label, ms, n * bpr / ms / 1e6, int(eng.totals[0])))


if __name__ == "__main__":
    main()
