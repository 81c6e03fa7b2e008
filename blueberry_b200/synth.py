"""Synthetic Hi-C contact records of the shapes BASELINE.json names (host / numpy version).

Every bin of every chromosome is a fragment (mid = i*R + R//2), and EVERY pair
(i, i+d), 0 <= d <= K, is an explicit record (zeros included) in row-major order,
so the record count equals the named shape.  Counts are Poisson with mean
``A * b_i * b_j * (d+1)^-1.08 * (1 + 4*loop_ij)``: a distance decay, a log-normal
per-bin visibility ``b`` (the ICE bias the significance pass corrects for) and a
sparse set of 5x-enriched "loops" that gives a real significant tail.

The device generator used by bench.py (csrc/synth.cu) draws from the same model
with a counter-based hash RNG; the two are NOT bit-identical and do not need to
be: parity tests upload the host arrays, the bench copies a sample of the device
arrays back for the CPU baseline.
"""
import numpy as np

HG19_LENGTHS = [
    249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663, 146364022,
    141213431, 135534747, 135006516, 133851895, 115169878, 107349540, 102531392, 90354753,
    81195210, 78077248, 59128983, 63025520, 48129895, 51304566, 155270560,
]


def n_bins_of(length, resolution):
    return -(-int(length) // int(resolution))


def n_pairs_of(n_bins, K):
    """Records of one chromosome: all (i, i+d) with 0 <= d <= K and i+d < n_bins."""
    K = min(int(K), n_bins - 1)
    return (K + 1) * n_bins - K * (K + 1) // 2


def make_fragments(chrom_bins, resolution):
    """(frag_chrom int32, frag_mid int32) for chromosomes with the given bin counts."""
    chrom = np.concatenate([np.full(n, c, dtype=np.int32) for c, n in enumerate(chrom_bins)])
    mid = np.concatenate([np.arange(n, dtype=np.int64) * resolution + resolution // 2 for n in chrom_bins])
    return chrom, mid.astype(np.int32)


def make_bias(chrom_bins, seed, sigma=0.25):
    """Per-bin visibility b ~ LogNormal(0, sigma); about 1% falls outside [0.5, 2]."""
    rng = np.random.default_rng(seed + 7919)
    return [np.exp(rng.normal(0.0, sigma, size=n)) for n in chrom_bins]


def make_contacts(chrom_bins, resolution, max_dist, depth, seed=20161108, bias=None, loop_rate=1e-3,
                  keep_zeros=True):
    """Return dict(chrom, mid1, mid2, count) int32 arrays, row-major per chromosome.

    depth: the scale A of the Poisson mean (mean count of a d=0 pair with b=1).
    """
    rng = np.random.default_rng(seed)
    K = int(max_dist) // int(resolution)
    out = {"chrom": [], "mid1": [], "mid2": [], "count": []}
    for c, n in enumerate(chrom_bins):
        Kc = min(K, n - 1)
        i = np.repeat(np.arange(n, dtype=np.int64), Kc + 1)
        d = np.tile(np.arange(Kc + 1, dtype=np.int64), n)
        ok = i + d < n
        i, d = i[ok], d[ok]
        lam = depth * (d + 1.0) ** -1.08
        if bias is not None:
            lam = lam * bias[c][i] * bias[c][i + d]
        loop = (rng.random(i.size) < loop_rate) & (d >= 5)
        lam = lam * (1.0 + 4.0 * loop)
        cnt = rng.poisson(lam).astype(np.int64)
        if not keep_zeros:
            nz = cnt > 0
            i, d, cnt = i[nz], d[nz], cnt[nz]
        out["chrom"].append(np.full(i.size, c, dtype=np.int32))
        out["mid1"].append((i * resolution + resolution // 2).astype(np.int32))
        out["mid2"].append(((i + d) * resolution + resolution // 2).astype(np.int32))
        out["count"].append(cnt.astype(np.int32))
    return {k: np.concatenate(v) for k, v in out.items()}
