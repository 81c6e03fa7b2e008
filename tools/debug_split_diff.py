"""Debug helper (GPU): where does the split K4 differ from the direct kernel?  python tools/debug_split_diff.py [seed]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_genome_pass import _random_shards, _direct
from blueberry_b200.distributed import GenomePass

for seed in [int(a) for a in sys.argv[1:]] or [1, 2]:
    dev = torch.device("cuda", 0)
    eng, shards = _random_shards(seed, dev, seed >= 8)
    gp = GenomePass(eng, group=False, q_values=True)
    gp.attach(shards)
    fit = gp.run()
    pn = gp.p.cpu().numpy()[:gp.rows]
    po, qo, starts = _direct(eng, shards, dev)
    po = po[:gp.rows]
    nan_diff = np.isnan(pn) != np.isnan(po)
    both = ~np.isnan(pn) & ~np.isnan(po)
    val_diff = both & (pn != po)
    print("seed", seed, "rows", gp.rows, "bias", eng.bias is not None, "S", fit.S, "exact", gp.last_score.exact,
          "nan-pattern diffs", int(nan_diff.sum()), "value diffs", int(val_diff.sum()))
    cnt = np.concatenate([np.pad(s.count.cpu().numpy(), (0, (-s.n) % 4)) for s in shards])
    m1 = np.concatenate([np.pad(s.mid1.cpu().numpy(), (0, (-s.n) % 4)) for s in shards])
    m2 = np.concatenate([np.pad(s.mid2.cpu().numpy(), (0, (-s.n) % 4)) for s in shards])
    for i in np.flatnonzero(nan_diff)[:8]:
        print("  nan row", i, "cnt", cnt[i], "d", m2[i] - m1[i], "new", pn[i], "old", po[i])
    for i in np.flatnonzero(val_diff)[:12]:
        print("  val row", i, "cnt", cnt[i], "d", m2[i] - m1[i], "new %.17g old %.17g rel %.3g" % (pn[i], po[i], pn[i] / po[i] - 1))
    if val_diff.any():
        c = cnt[val_diff]
        print("  counts of differing rows:", np.unique(c, return_counts=True))
