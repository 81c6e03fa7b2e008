"""Files in -> files out: FitHiC.fit_transform on synthetic gzip text, with the time of each leg.

    python tools/bench_files.py [bins]       (needs a GPU; default 2500 bins at 10 kb, every pair: 3.1 M rows)

The reference does this at ~3e4 rows/s (per-line Python over gzip text).
"""
import gzip
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from blueberry_b200 import _io, _lib, fithic


def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
    R, K = 10000, nb - 1
    lib = _lib.load()
    dev = torch.device("cuda:0")
    P = int(lib.bbk_synth_n_pairs(nb, K))
    rng = np.random.default_rng(3)
    bias = np.exp(rng.normal(0.0, 0.25, size=nb))
    bias_dev = torch.from_numpy(bias).to(dev)
    cols = [torch.empty(P, dtype=torch.int32, device=dev) for _ in range(3)]
    _lib.check(lib.bbk_synth_contacts(nb, K, R, 2000.0, 1.08, 11, _lib.ptr(bias_dev), _lib.ptr(cols[0]), _lib.ptr(cols[1]),
                                      _lib.ptr(cols[2]), _lib.stream_ptr()), "synth")
    m1, m2, c = (t.cpu().numpy() for t in cols)
    tmp = tempfile.mkdtemp(prefix="bbk_files_")
    inter, frags, biases = (os.path.join(tmp, n) for n in ("interactions.gz", "fragments.gz", "biases.gz"))
    import pandas as pd
    t = time.time()
    pd.DataFrame({"a": "chr21", "b": m1, "c": "chr21", "d": m2, "e": c}).to_csv(inter, sep="\t", header=False, index=False,
                                                                              compression={"method": "gzip", "compresslevel": 1})
    mids = np.arange(nb, dtype=np.int64) * R + R // 2
    with gzip.open(frags, "wt") as fh:
        fh.write("".join("chr21\t%d\t0\t0\t0\n" % m for m in mids))
    with gzip.open(biases, "wt") as fh:
        fh.write("".join("chr21\t%d\t%r\n" % (m, float(b)) for m, b in zip(mids, bias)))
    print("inputs written (not timed as part of the pass): %d rows, %.1f MB gz, %.1f s" % (P, os.path.getsize(inter) / 1e6, time.time() - t))
    legs = {}
    real_parse, real_write = fithic._parse_interactions, fithic._write_significances

    def timed(name, fn):
        def wrap(*a, **k):
            t0 = time.time()
            r = fn(*a, **k)
            legs[name] = legs.get(name, 0.0) + time.time() - t0
            return r
        return wrap
    fithic._parse_interactions = timed("parse interactions", real_parse)
    fithic._write_significances = timed("write significances", real_write)
    model = fithic.FitHiC(os.path.join(tmp, "lib"), R, n_bins=100, max_dist=K * R)
    model.fit_transform(inter, frags, biases)                 # warm-up (CUDA context, tables)
    legs.clear()
    t0 = time.time()
    model.fit_transform(inter, frags, biases)
    total = time.time() - t0
    out = os.path.join(tmp, "lib.spline_pass1.res%d.significances.txt.gz" % R)
    rest = total - sum(legs.values())
    print("fit_transform: %.2f s total = %.2f M rows/s | parse %.2f s, write %.2f s (%.1f MB), everything else incl. the GPU pass %.2f s"
          % (total, P / total / 1e6, legs.get("parse interactions", 0), legs.get("write significances", 0), os.path.getsize(out) / 1e6, rest))
    print("reference pace (SURVEY.md section 6): ~3e4 rows/s -> %.0f s for this input" % (P / 3e4))


if __name__ == "__main__":
    main()
