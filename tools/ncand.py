"""Scratch: candidate count of the q-value step on the bench workload (reads BhState from the workspace)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from blueberry_b200 import _lib
from blueberry_b200.engine import BiasTables, PassEngine, Shard
lib = _lib.load(); dev = torch.device("cuda:0")
R, nb, K = 5000, 49851, 2000
P = int(lib.bbk_synth_n_pairs(nb, K))
rng = np.random.default_rng(20240 + 0)
import bench
rng = np.random.default_rng(bench.SEED)
bias_host = np.exp(rng.normal(0.0, 0.25, size=nb)); bias_dev = torch.from_numpy(bias_host).to(dev)
cols = [torch.empty(P, dtype=torch.int32, device=dev) for _ in range(3)]
_lib.check(lib.bbk_synth_contacts(nb, K, R, bench.DEPTH, bench.DECAY, bench.SEED, _lib.ptr(bias_dev), _lib.ptr(cols[0]), _lib.ptr(cols[1]), _lib.ptr(cols[2]), _lib.stream_ptr()), "synth")
eng = PassEngine(R, 100, 0, K * R, nb, dev); eng.set_fragments([nb], [(nb - 1) * R])
eng.set_bias(BiasTables([np.where((bias_host < 0.5) | (bias_host > 2), -1.0, bias_host)], [R // 2], dev))
p = torch.empty(P + 1, dtype=torch.float64, device=dev)[:P]; q = torch.empty(P + 1, dtype=torch.float64, device=dev)[:P]
eng.run([Shard(*cols)], [p], [q])
torch.cuda.synchronize()
st = eng.bh_ws[:64].cpu().numpy().view(np.uint64)
print("n_cand", st[0], "n_ones", st[1], "n_nan", st[2], "n_valid", st[3], "tau_key", hex(int(st[4])))
print("skip", eng.bh_ws[64:64+40].cpu().numpy().view(np.int32))
print("q<1:", int((q < 1).sum()), "q<=0.01:", int((q <= 0.01).sum()))
