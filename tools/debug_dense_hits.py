import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import test_gpu_genome_pass as T
from blueberry_b200 import synth
from blueberry_b200.distributed import GenomePass
from blueberry_b200.engine import BiasTables, PassEngine, Shard
dev = torch.device("cuda", 0)
R, bins, max_dist = 5000, [700, 400], 600 * 5000
bias = synth.make_bias(bins, 8, sigma=0.3)
c = synth.make_contacts(bins, R, max_dist, 300.0, 3, bias)
chrom, m1, m2, cn = c["chrom"], c["mid1"].copy(), c["mid2"].copy(), c["count"].copy()
eng = PassEngine(R, 100, 0, max_dist, max(bins), dev)
eng.set_fragments(bins, [(b - 1) * R for b in bins])
tabs = [np.where((b < 0.5) | (b > 2), -1.0, b) for b in bias]
eng.set_bias(BiasTables(tabs, [R // 2, R // 2], dev))
sel0, sel1 = np.flatnonzero(chrom == 0), np.flatnonzero(chrom == 1)
cn[sel0[:5000]] += 900
shards = [Shard(T._t32(m1[ix], dev), T._t32(m2[ix], dev), T._t32(cn[ix], dev), chrom=ci) for ci, ix in ((0, sel0), (1, sel1))]
gp = GenomePass(eng, group=False, q_values=True); gp.attach(shards); gp.run()
p_old, q_old, _ = T._direct(eng, shards, dev)
a, b = gp.p.cpu().numpy()[:gp.rows], p_old[:gp.rows]
print("nan pattern same", np.array_equal(np.isnan(a), np.isnan(b)))
ok = ~np.isnan(a)
print("ones same", np.array_equal(a[ok] == 1.0, b[ok] == 1.0), "zeros same", np.array_equal(a[ok] == 0.0, b[ok] == 0.0))
pos = ok & (a > 0) & (b > 0)
rel = np.abs(a[pos] / b[pos] - 1)
idx = np.flatnonzero(pos)[np.argsort(-rel)[:8]]
cnt_all = np.concatenate([cn[sel0], np.zeros((-len(sel0)) % 4, dtype=cn.dtype), cn[sel1]])
for i in idx:
    print(i, cnt_all[i] if i < len(cnt_all) else None, repr(a[i]), repr(b[i]), a[i] / b[i] - 1)
z = np.flatnonzero(ok & ((a == 0) != (b == 0)))[:5]
for i in z: print("zero mismatch", i, repr(a[i]), repr(b[i]))
