"""Times the UNMODIFIED reference (blueberry/fithic.py, patched in memory for Python 3 by oracle/ref_loader.py: a closed list of
syntax edits) end to end on gzip text, one core - BASELINE.md section 3, legs A/B.  Build container only (/root/reference is
absent on the GPU box); the result is committed as profiles/true_reference_timing.json and quoted by bench.py's reference arm.

    python tools/time_true_reference.py [rows]
"""
import json
import os
import platform
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from blueberry_b200 import synth
from oracle import ref_loader, run_reference


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    if not ref_loader.reference_available():
        raise SystemExit("needs /root/reference")
    # a prefix of BASELINE config 1's shape: chr21 @ 10 kb (4,813 bins), no effective distance cap, zeros kept
    R, nb = 10000, 4813
    n_bins_used = 1
    while n_bins_used * (n_bins_used + 1) // 2 < rows:
        n_bins_used += 1
    bins = [n_bins_used]
    bias = synth.make_bias(bins, 21)
    fc, fm = synth.make_fragments(bins, R)
    c = synth.make_contacts(bins, R, 10 ** 9, 3000.0, 2021, bias)
    n = len(c["count"])
    tmp = tempfile.mkdtemp(prefix="bbk_trueref_")
    ref = ref_loader.load_reference_fithic()
    inter, frags, biasf = (os.path.join(tmp, f) for f in ("interactions.gz", "fragments.gz", "biases.gz"))
    run_reference.write_interactions(inter, c["chrom"], c["mid1"], c["chrom"], c["mid2"], c["count"])
    run_reference.write_fragments(frags, fc, fm)
    run_reference.write_biases(biasf, np.zeros(bins[0], dtype=np.int32), fm, bias[0])
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        model = ref.FitHiC("lib", R, n_bins=100)
        t0 = time.perf_counter()
        model.fit_transform(inter, frags, biasf)                 # fithic.py:85-108: gz text in, significances.txt.gz out
        wall = time.perf_counter() - t0
    finally:
        os.chdir(cwd)
    out = {"what": "reference FitHiC.fit_transform (fithic.py:85-108), gz text in / gz text out, single thread (the reference is "
                   "single-threaded)", "rows": n, "bins": bins[0], "resolution": R, "seconds": wall, "pairs_per_sec": n / wall,
           "cores_used": 1, "cores_online": os.cpu_count(), "host": platform.processor() or platform.machine(),
           "python": platform.python_version(), "where": "build container (no GPU); /root/reference is absent on the GPU box"}
    print(json.dumps(out))
    with open(os.path.join(ROOT, "profiles", "true_reference_timing.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
