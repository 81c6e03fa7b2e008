"""Mint tests/golden/*.npz by EXECUTING the reference (build container only).

    python -m oracle.make_golden

The reference ships no fixtures, so these are outputs of its own code
(fithic.py patched in memory per ref_loader.py; blueberry.pyx compiled verbatim)
run in this image: python 3.12.3, numpy 2.3.5, scipy 1.18.1, scikit-learn 1.9.0.
Inputs are stored beside the outputs so the GPU box needs nothing but the .npz.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from blueberry_b200 import synth            # noqa: E402
from oracle import ref_loader, run_reference  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _pass_case(name, bins, R, max_dist, min_dist, depth, seed, with_bias, keep_zeros, n_bins=100, messy=False):
    bias = synth.make_bias(bins, seed) if with_bias else None
    fc, fm = synth.make_fragments(bins, R)
    c = synth.make_contacts(bins, R, max_dist if max_dist > 0 else 10**9, depth, seed, bias, keep_zeros=keep_zeros)
    chr1, chr2 = c["chrom"].copy(), c["chrom"].copy()
    mid1, mid2, cnt = c["mid1"].copy(), c["mid2"].copy(), c["count"].copy()
    if messy:
        rng = np.random.default_rng(seed + 1)
        n = len(cnt)
        # inter-chromosomal rows whose midpoint difference is in range (fithic.py:247,256 never checks chr)
        k = rng.choice(n, 200, replace=False)
        chr2[k] = (chr2[k] + 1) % len(bins)
        # off-grid distances (counted in S, not in mainDic: fithic.py:260-263), negative distances, duplicates
        k = rng.choice(n, 100, replace=False)
        mid2[k] += 1234
        k = rng.choice(n, 50, replace=False)
        mid1[k], mid2[k] = mid2[k].copy(), mid1[k].copy()
        dup = rng.choice(n, 300, replace=False)
        chr1, chr2 = np.concatenate([chr1, chr1[dup]]), np.concatenate([chr2, chr2[dup]])
        mid1, mid2, cnt = np.concatenate([mid1, mid1[dup]]), np.concatenate([mid2, mid2[dup]]), np.concatenate([cnt, cnt[dup]])
        perm = rng.permutation(len(cnt))
        chr1, chr2, mid1, mid2, cnt = chr1[perm], chr2[perm], mid1[perm], mid2[perm], cnt[perm]
    barr = None
    if with_bias:
        bc = np.concatenate([np.full(b, i, dtype=np.int32) for i, b in enumerate(bins)])
        bv = np.concatenate(bias)
        bm = fm.copy()
        if messy:  # a duplicated locus (first occurrence wins, fithic.py:153) and a missing one (default 1.0, :418-425)
            bc, bm, bv = np.concatenate([bc[:1], bc[:-1]]), np.concatenate([bm[:1], bm[:-1]]), np.concatenate([[1.5], bv[:-1]])
        barr = (bc, bm, bv)
    ref = run_reference.run_reference_pass(fc, fm, chr1, mid1, chr2, mid2, cnt, R, n_bins=n_bins,
                                           min_dist=min_dist, max_dist=max_dist, bias=barr)
    out = ref["out"]
    save = dict(
        resolution=R, n_bins=n_bins, min_dist_arg=min_dist, max_dist_arg=max_dist,
        frag_chrom=fc, frag_mid=fm, chr1=chr1, mid1=mid1, chr2=chr2, mid2=mid2, count=cnt,
        has_bias=with_bias,
        bias_chrom=barr[0] if barr else np.zeros(0, np.int32),
        bias_mid=barr[1] if barr else np.zeros(0, np.int32),
        bias_val=barr[2] if barr else np.zeros(0),
        ref_possible=ref["possible"], ref_observed=ref["observed"], ref_x=ref["x"], ref_y=ref["y"],
        ref_spline_x=ref["spline_x"], ref_spline_y=ref["spline_y"], ref_residual=ref["residual"],
        ref_out_chr1=out["chr1"], ref_out_mid1=out["mid1"], ref_out_chr2=out["chr2"], ref_out_mid2=out["mid2"],
        ref_out_count=out["count"], ref_out_p=out["p"], ref_out_q=out["q"],
    )
    for k in ("S", "intra_in_range_count", "intra_all_sum", "intra_all_count", "inter_all_sum", "inter_all_count",
              "min_obs_dist", "max_obs_dist", "max_possible_dist", "possible_intra_in_range",
              "possible_intra_all", "possible_inter_all", "min_dist", "max_dist"):
        save["ref_" + k] = np.int64(ref[k])
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **save)
    print(name, "pairs", len(cnt), "kept", len(out["p"]), "bins", len(ref["x"]), "S", ref["S"])


def _bh_case():
    ref = ref_loader.load_reference_fithic()
    cy = ref_loader.load_reference_cython()
    rng = np.random.default_rng(11)
    cases = {}
    p = rng.random(1000)
    p[rng.choice(1000, 200, replace=False)] = 1.0          # the count==0 rows
    p[rng.choice(1000, 50, replace=False)] = p[0]          # ties
    p[:5] = [0.0, 1e-300, 5e-324, 1.0, 0.5]
    for tag, arr, n in (("a", p, 1000), ("b", p, 25000), ("c", p[:7], 3),
                        ("doc", np.array([0.03, 0.4, 0.7, 0.01]), 10)):     # fithic.py:460 usage comment
        cases["p_" + tag] = arr
        cases["n_" + tag] = np.int64(n)
        cases["q_py_" + tag] = np.array(ref.benjamini_hochberg_correction(list(arr), n), dtype=np.float64)
        srt = np.sort(arr)
        cases["q_cy_sorted_" + tag] = np.asarray(cy.benjamini_hochberg(srt, n), dtype=np.float64)
    reg = np.sort(rng.choice(40000, 3000, replace=False)).astype(np.float64) * 5000 + 2500
    cases["regions_sorted"] = reg
    cases["band_sorted"] = np.int64(cy.count_band_regions(reg))
    shuf = reg[rng.permutation(len(reg))][:1500].copy()
    cases["regions_shuffled"] = shuf
    cases["band_shuffled"] = np.int64(cy.count_band_regions(shuf))
    np.savez_compressed(os.path.join(GOLDEN, "bh_band.npz"), **cases)
    print("bh_band", {k: v.shape for k, v in cases.items() if k.startswith("q_py")}, cases["band_sorted"], cases["band_shuffled"])


def _pass2_case():
    """Second pass on the inputs of pass_bias_dense (outputs only; inputs are in pass_bias_dense.npz)."""
    from oracle import fithic_oracle as fo
    g = np.load(os.path.join(GOLDEN, "pass_bias_dense.npz"))
    R = int(g["resolution"])
    bd, _ = fo.read_bias_arrays(g["bias_chrom"], g["bias_mid"], g["bias_val"])
    o = fo.fithic_arrays(g["frag_chrom"], g["frag_mid"], g["chr1"], g["mid1"], g["chr2"], g["mid2"], g["count"], R,
                         int(g["n_bins"]), int(g["ref_min_dist"]), int(g["ref_max_dist"]), bias=bd)
    ref2 = run_reference.run_reference_two_pass(g["frag_chrom"], g["frag_mid"], g["chr1"], g["mid1"], g["chr2"], g["mid2"], g["count"],
                                                R, o.keep, n_bins=int(g["n_bins"]), min_dist=int(g["min_dist_arg"]),
                                                max_dist=int(g["max_dist_arg"]), bias=(g["bias_chrom"], g["bias_mid"], g["bias_val"]))
    out = ref2["out"]
    np.savez_compressed(os.path.join(GOLDEN, "pass2_bias_dense.npz"), threshold=ref2["threshold"], n_outliers=ref2["n_outliers"],
                        outlier_idx=ref2["outlier_idx"], ref2_observed=ref2["observed"], ref2_S=np.int64(ref2["S"]), ref2_x=ref2["x"],
                        ref2_y=ref2["y"], ref2_spline_x=ref2["spline_x"], ref2_spline_y=ref2["spline_y"], ref2_out_mid1=out["mid1"],
                        ref2_out_mid2=out["mid2"], ref2_out_count=out["count"], ref2_out_p=out["p"])
    print("pass2_bias_dense outliers", ref2["n_outliers"], "S", ref2["S"], "rows", len(out["p"]))


def _decimate_case():
    """FithicContactMap.decimate (datatypes.pyx:317-339) run from the reference's own method source."""
    rng = np.random.default_rng(17)
    n = 6000
    m1 = rng.integers(0, 500, n) * 1000 + 500
    m2 = m1 + rng.integers(0, 80, n) * 1000
    mp = np.stack([m1, m2, rng.integers(0, 60, n), rng.random(n) ** 4, np.minimum(rng.random(n) ** 2 * 3, 1.0)], axis=1).astype(np.float64)
    mp[rng.choice(n, 200, replace=False), 4] = -1.0          # the literal q the reference's own files carry (fithic.py:435)
    mp = np.concatenate([mp, mp[rng.choice(n, 500)]])        # exact duplicates
    mp = mp[rng.permutation(len(mp))]                        # not sorted: groups appear interleaved
    out5 = run_reference.run_reference_decimate(mp, 5000)
    out25 = run_reference.run_reference_decimate(out5, 25000)
    np.savez_compressed(os.path.join(GOLDEN, "decimate.npz"), map_in=mp, ref_5000=out5, ref_25000_of_5000=out25)
    print("decimate", mp.shape, "->", out5.shape, "->", out25.shape)


def _contact_map_case():
    """ContactMap.__init__ + normalize (datatypes.pyx:88-171), the reference compiled verbatim."""
    rng = np.random.default_rng(29)
    nb, R = 120, 5000
    kr = rng.random(nb) + 0.5
    kr[rng.choice(nb, 6, replace=False)] = np.nan                   # unmappable bins, as in real KRnorm files
    ke = np.sort(rng.random(nb) * 80 + 0.5)[::-1].copy()
    pairs = [(i, j) for i in range(nb + 1) for j in range(i, min(i + 40, nb + 1))]
    sel = rng.choice(len(pairs), 2500, replace=False)
    b1 = np.array([pairs[k][0] for k in sel])
    b2 = np.array([pairs[k][1] for k in sel])
    swap = rng.random(len(sel)) < 0.3                               # rows of the lower triangle
    pos1 = np.where(swap, b2, b1) * R
    pos2 = np.where(swap, b1, b2) * R
    cnt = rng.integers(1, 500, len(sel)).astype(np.float64)
    cnt[rng.choice(len(sel), 10, replace=False)] = np.nan           # nan_to_num at ingest (:104)
    before, after, regions, n_bins = run_reference.run_reference_contact_map(pos1, pos2, cnt, kr, ke, R)
    np.savez_compressed(os.path.join(GOLDEN, "contact_map.npz"), resolution=R, pos1=pos1, pos2=pos2, count=cnt, kr_norm=kr,
                        kr_expected=ke, ref_matrix=before, ref_normalized=after, ref_regions=regions, ref_n_bins=n_bins)
    print("contact_map", before.shape, "records", len(sel), "nonzero after", int(np.count_nonzero(after)))


def _extract_case():
    """utils.extract_contacts (utils.py:31-90) run from the reference's own function source on two chromosomes, and the
    genome-wide q-values composed from it as in SURVEY.md 3.2 with the reference's compiled benjamini_hochberg."""
    rng = np.random.default_rng(41)
    bb = ref_loader.load_reference_cython()
    save, parts, n_total = {}, [], 0
    for chrom, n in ((3, 5000), (11, 3500)):
        m1 = rng.integers(0, 3000, n) * 5000 + 2500
        m2 = m1 + rng.integers(0, 2400, n) * 5000                      # distances 0 .. 12 Mb: both ends of the band are crossed
        p = rng.random(n) ** 6
        p[rng.choice(n, 40, replace=False)] = 1.0
        p[:30] = p[30:60]                                              # tied p-values
        mp = np.stack([m1, m2, rng.integers(1, 90, n), p, np.full(n, -1.0)], axis=1).astype(np.float64)
        run_reference.write_reference_significances(mp, chrom, 5000)
        mp = run_reference.reference_map_as_read(chrom, 5000)          # what the reference holds after ITS read of the file
        contact, band = run_reference.run_reference_extract_contacts(chrom, 5000, alpha=0.2, n_regions=True)
        contact_all = run_reference.run_reference_extract_contacts(chrom, 5000)
        save["map_%d" % chrom], save["ref_contacts_%d" % chrom], save["ref_band_%d" % chrom] = mp, contact, np.int64(band)
        save["ref_contacts_noalpha_%d" % chrom] = contact_all
        parts.append(contact)
        n_total += int(band)
    allc = np.concatenate(parts)
    order = np.argsort(allc[:, 4], kind="stable")
    q = np.empty(len(order))
    q[order] = np.asarray(bb.benjamini_hochberg(np.ascontiguousarray(allc[order, 4]), n_total))
    np.savez_compressed(os.path.join(GOLDEN, "extract_contacts.npz"), alpha=0.2, chroms=np.array([3, 11]), ref_q=q, ref_n=np.int64(n_total), **save)
    print("extract_contacts", [len(x) for x in parts], "band total", n_total, "q<=0.01:", int((q <= 0.01).sum()))


def main():
    if not ref_loader.reference_available():
        raise SystemExit("needs /root/reference (build container only)")
    os.makedirs(GOLDEN, exist_ok=True)
    # with ICE biases, zeros kept, two chromosomes (the shape of the BASELINE configs, shrunk)
    _pass_case("pass_bias_dense", [260, 170], 10000, 1500000, -1, 150.0, 5, True, True)
    # no bias file, zeros dropped (what real Fit-Hi-C input looks like), default 10 Mb cap larger than the chromosome
    _pass_case("pass_nobias_sparse", [900], 5000, -1, -1, 8.0, 6, False, False)
    # everything the reference tolerates: inter rows, off-grid / negative distances, duplicates, shuffled order,
    # min_dist > 0, bias file with a duplicate and a missing locus, n_bins != 100
    _pass_case("pass_messy", [300, 220, 90], 5000, 1000000, 20000, 60.0, 7, True, True, n_bins=40, messy=True)
    _bh_case()
    _pass2_case()
    _decimate_case()
    _contact_map_case()
    _extract_case()


if __name__ == "__main__":
    main()
