"""world_size-2 gloo run (CPU): the only data-path exchange of the pass - the all-reduce of K1's
per-distance table and totals - gives every rank the table a single process would have computed."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from blueberry_b200 import sharding, synth
    from blueberry_b200.engine import reduce_distance_stats
    from oracle import fithic_oracle as fo
    R, bins, max_dist = 10000, [220, 180, 90, 60], 1200000
    fc, fm = synth.make_fragments(bins, R)
    c = synth.make_contacts(bins, R, max_dist, 30.0, 17)
    loads = [synth.n_pairs_of(b, max_dist // R) for b in bins]
    owner = np.array(sharding.lpt_assign(loads, world))
    frag = fo.generate_frag_pairs(fc, fm, R, 0, max_dist)
    nkeys = len(frag.possible)
    mine = owner[c["chrom"]] == rank                                   # this rank's chromosomes
    st = fo.read_interactions(nkeys, R, c["chrom"][mine], c["mid1"][mine], c["chrom"][mine], c["mid2"][mine],
                              c["count"][mine], 0, max_dist)
    obs = torch.from_numpy(st.observed.copy())
    totals = torch.tensor([st.S, st.intra_in_range_count, st.intra_all_sum, st.intra_all_count, st.inter_all_sum,
                           st.inter_all_count, st.min_obs_dist, st.max_obs_dist], dtype=torch.int64)
    reduce_distance_stats(obs, totals)
    full = fo.read_interactions(nkeys, R, c["chrom"], c["mid1"], c["chrom"], c["mid2"], c["count"], 0, max_dist)
    ok = bool(np.array_equal(obs.numpy(), full.observed)) and totals.tolist() == [
        full.S, full.intra_in_range_count, full.intra_all_sum, full.intra_all_count, full.inter_all_sum,
        full.inter_all_count, full.min_obs_dist, full.max_obs_dist]
    # the fit that follows is a pure function of the reduced table: identical on both ranks
    x, y, _, _ = fo.calculate_probabilities(frag.possible, obs.numpy(), int(totals[0]), 50, R, 0, max_dist)
    q.put((rank, ok, int(mine.sum()), float(np.sum(y)), float(np.sum(x))))
    dist.destroy_process_group()


def test_allreduce_of_distance_stats_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    assert res[0][2] > 0 and res[1][2] > 0                              # both ranks held records
    assert res[0][3:] == res[1][3:]                                     # identical bins on both ranks
