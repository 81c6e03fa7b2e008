"""Run under torchrun on N GPUs (not collected by pytest): genome-wide q-values across ranks and the all-reduced
distance table against the single-process oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    from blueberry_b200 import sharding, synth
    from blueberry_b200.engine import BiasTables, PassEngine, Shard
    from oracle import fithic_oracle as fo

    R, bins, max_dist = 10000, [420, 380, 300, 260, 150, 90][: max(world + 2, 4)], 2_000_000
    fc, fm = synth.make_fragments(bins, R)
    bias = synth.make_bias(bins, 4)
    c = synth.make_contacts(bins, R, max_dist, 90.0, 23, bias)                 # same on every rank (seeded)
    owner = np.array(sharding.lpt_assign([synth.n_pairs_of(b, max_dist // R) for b in bins], world))
    mine = owner[c["chrom"]] == rank
    t32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
    shard = Shard(t32(c["mid1"][mine]), t32(c["mid2"][mine]), t32(c["count"][mine]), t32(c["chrom"][mine]), t32(c["chrom"][mine]))
    info_n = bins
    nkeys = (max(bins) - 1) + 1
    eng = PassEngine(R, 100, 0, max_dist, nkeys, dev)
    eng.set_fragments(info_n, [(b - 1) * R for b in bins])
    tabs = [np.where((b < 0.5) | (b > 2), -1.0, b) for b in bias]
    eng.set_bias(BiasTables(tabs, [R // 2] * len(bins), dev))
    n = shard.n
    p = torch.empty((n + 1) & ~1, dtype=torch.float64, device=dev)[:n]
    q = torch.empty((n + 1) & ~1, dtype=torch.float64, device=dev)[:n]
    eng.hist([shard])
    eng.allreduce_stats()
    eng.fit()
    eng.p_hist.zero_()
    eng.pvalues(shard, p, with_hist=True)
    n_all = eng.qvalues_global(p, q, hist=eng.p_hist)
    torch.cuda.synchronize()
    eng.read_fit()

    # single-process oracle on ALL records
    bc = np.concatenate([np.full(b, i) for i, b in enumerate(bins)])
    bd, _ = fo.read_bias_arrays(bc, fm, np.concatenate(bias))
    ref = fo.fithic_arrays(fc, fm, c["chrom"], c["mid1"], c["chrom"], c["mid2"], c["count"], R, 100, 0, max_dist, bias=bd)
    ok = np.array_equal(eng.obs_sum.cpu().numpy(), ref.contacts.observed) and int(eng.totals[0].item()) == ref.contacts.S
    pg = p.cpu().numpy()
    keep_ref = ref.keep[mine]
    ok &= np.array_equal(pg <= 1, keep_ref)
    kk = keep_ref & (ref.p[mine] > 0)
    err = np.abs(np.log10(pg[kk]) - np.log10(ref.p[mine][kk])).max()
    ok &= err <= 1e-6
    # genome-wide q: gather every rank's p, rank them together with the oracle, compare this rank's rows bit for bit
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n], dtype=torch.int64, device=dev))
    sizes = [int(s.item()) for s in sizes]
    cap = max(sizes)
    buf = torch.full((cap,), float("nan"), dtype=torch.float64, device=dev)
    buf[:n] = p
    allp = [torch.empty(cap, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(allp, buf)
    p_all = np.concatenate([a[:s].cpu().numpy() for a, s in zip(allp, sizes)])
    valid = ~np.isnan(p_all)
    q_ref_all = np.full(len(p_all), np.nan)
    q_ref_all[valid] = fo.benjamini_hochberg_correction(p_all[valid], int(valid.sum()))
    off = sum(sizes[:rank])
    q_mine = q.cpu().numpy()
    same = np.array_equal(np.isnan(q_mine), np.isnan(q_ref_all[off:off + n])) and \
        np.array_equal(q_mine[~np.isnan(q_mine)], q_ref_all[off:off + n][~np.isnan(q_mine)])
    ok &= same
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        sys.stderr.write("multi_gpu_check world=%d: records/rank %s, candidates gathered %d, max |dlog10 p| %.3g, q identical %s -> %s\n"
                         % (world, sizes, n_all, err, same, "OK" if flag.item() else "FAILED"))
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
