"""Debug aid: speculative vs exact mode of the streaming K4 on one random shard set (p, q, p histogram, score state)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_genome_pass import _random_shards
from blueberry_b200.distributed import GenomePass

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
eng, shards = _random_shards(seed, dev)
gp = GenomePass(eng, group=False, q_values=True)
gp.attach(shards)
res = []
for mode in (False, True, False, True):
    gp.force_exact = mode
    gp.p.fill_(7.0); gp.q.fill_(7.0)
    gp.run()
    torch.cuda.synchronize()
    sc = gp.last_score
    res.append((gp.p.cpu().numpy().copy(), gp.q.cpu().numpy().copy(), eng.p_hist.cpu().numpy().copy(), (sc.n_list, sc.n_cand, sc.overflow, sc.cand_overflow, sc.exact)))
    print("mode", mode, "state", res[-1][3], "hist ones/nans", res[-1][2][4096], res[-1][2][4097], "hist sum", res[-1][2][:4096].sum())
for a, b in ((0, 1), (0, 2), (1, 3)):
    pa, qa, ha, _ = res[a]; pb, qb, hb, _ = res[b]
    dp = ~((pa == pb) | (np.isnan(pa) & np.isnan(pb)))
    dq = ~((qa == qb) | (np.isnan(qa) & np.isnan(qb)))
    print("runs", a, b, "p diffs", int(dp.sum()), "q diffs", int(dq.sum()), "hist diffs", int((ha != hb).sum()))
    cnt = np.full(len(pa), -99, dtype=np.int64); m1 = cnt.copy(); m2 = cnt.copy()
    for sh, o in zip(gp.shards, gp.offsets):
        cnt[o:o + sh.n] = sh.count.cpu().numpy(); m1[o:o + sh.n] = sh.mid1.cpu().numpy(); m2[o:o + sh.n] = sh.mid2.cpu().numpy()
    for i in np.flatnonzero(dp)[:10]:
        print("   row", i, "count", cnt[i], "mids", m1[i], m2[i], "p", repr(pa[i]), repr(pb[i]), "rel", pa[i] / pb[i] - 1)
    for i in np.flatnonzero(dq)[:8]:
        print("   row", i, "p", pa[i], pb[i], "q", qa[i], qb[i])
    for i in np.flatnonzero(ha != hb)[:8]:
        print("   bin", i, ha[i], hb[i])
