"""Where the time of FitHiC.fit_transform_arrays goes (chr1 @ 5 kb, numpy columns in, numpy p / q out)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from blueberry_b200 import _lib, synth
from blueberry_b200.fithic import FitHiC
import blueberry_b200.fithic as F

dev = torch.device("cuda", 0)
lib = _lib.load()
R, K = 5000, 2000
nb = -(-249250621 // R)
n = int(lib.bbk_synth_n_pairs(nb, K))
rng = np.random.default_rng(1)
bias = np.exp(rng.normal(0.0, 0.25, size=nb))
cols = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(3)]
bdev = torch.from_numpy(bias).to(dev)
_lib.check(lib.bbk_synth_contacts_range(nb, K, R, 600.0, 1.08, 7, _lib.ptr(bdev), 0, n, _lib.ptr(cols[0]), _lib.ptr(cols[1]), _lib.ptr(cols[2]), _lib.stream_ptr()), "synth")
m1, m2, cn = (c.cpu().numpy() for c in cols)
del cols
fm = np.arange(nb, dtype=np.int64) * R + R // 2
fc = np.zeros(nb, dtype=np.int32)
model = FitHiC("t", R, n_bins=100, max_dist=10_000_000)
import cProfile, pstats
for it in range(3):
    t0 = time.perf_counter()
    res = model.fit_transform_arrays(None, m1, None, m2, cn, fc, fm, bias=(fc, fm, bias), q_values=True)
    print("call %d: %.1f ms" % (it, 1e3 * (time.perf_counter() - t0)))
pr = cProfile.Profile(); pr.enable()
res = model.fit_transform_arrays(None, m1, None, m2, cn, fc, fm, bias=(fc, fm, bias), q_values=True)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
