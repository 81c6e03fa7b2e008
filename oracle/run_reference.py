"""Drive the REAL reference (patched in memory, see ref_loader.py) on arrays.

Build container only (needs /root/reference).  TEST INFRASTRUCTURE ONLY.
"""
import gzip
import os
import tempfile

import numpy as np

from . import ref_loader


def _chrom_name(c):
    return "chr%d" % (int(c) + 1)


def write_interactions(path, chr1, mid1, chr2, mid2, count):
    """The reference's interactions format, fithic.py:245: chr1 mid1 chr2 mid2 count."""
    with gzip.open(path, "wt", compresslevel=1) as fh:
        for a, b, c, d, e in zip(chr1, mid1, chr2, mid2, count):
            fh.write("%s\t%d\t%s\t%d\t%d\n" % (_chrom_name(a), b, _chrom_name(c), d, e))


def write_fragments(path, frag_chrom, frag_mid):
    """fithic.py:288 reads the first two columns: chr mid."""
    with gzip.open(path, "wt", compresslevel=1) as fh:
        for c, m in zip(frag_chrom, frag_mid):
            fh.write("%s\t%d\t0\t0\t0\n" % (_chrom_name(c), m))


def write_biases(path, bias_chrom, bias_mid, bias_val):
    """fithic.py:143: chr mid bias."""
    with gzip.open(path, "wt", compresslevel=1) as fh:
        for c, m, b in zip(bias_chrom, bias_mid, bias_val):
            fh.write("%s\t%d\t%r\n" % (_chrom_name(c), m, float(b)))


def parse_significances(path):
    """Parse the reference's output (fithic.py:410-435) into arrays."""
    c1, m1, c2, m2, cnt, p, q = [], [], [], [], [], [], []
    with gzip.open(path, "rt") as fh:
        header = fh.readline()
        for line in fh:
            a, b, c, d, e, f, g = line.rstrip().split("\t")
            c1.append(int(a[3:]) - 1); m1.append(int(b)); c2.append(int(c[3:]) - 1); m2.append(int(d))
            cnt.append(int(e)); p.append(float(f)); q.append(float(g))
    return {"header": header, "chr1": np.array(c1, np.int32), "mid1": np.array(m1, np.int64),
            "chr2": np.array(c2, np.int32), "mid2": np.array(m2, np.int64),
            "count": np.array(cnt, np.int64), "p": np.array(p, np.float64), "q": np.array(q, np.float64)}


def run_reference_pass(frag_chrom, frag_mid, chr1, mid1, chr2, mid2, count, resolution,
                       n_bins=100, min_dist=-1, max_dist=-1, bias=None, workdir=None):
    """Run fithic.py's stages exactly as fithic() does (fithic.py:110-133), capturing intermediates.

    min_dist / max_dist use FitHiC.__init__'s sentinels (fithic.py:82-83).
    bias: None or (bias_chrom, bias_mid, bias_val).
    """
    ref = ref_loader.load_reference_fithic()
    tmp = workdir or tempfile.mkdtemp(prefix="bbk_ref_")
    inter = os.path.join(tmp, "interactions.gz")
    frags = os.path.join(tmp, "fragments.gz")
    write_interactions(inter, chr1, mid1, chr2, mid2, count)
    write_fragments(frags, frag_chrom, frag_mid)
    biasfile = "none"
    if bias is not None:
        biasfile = os.path.join(tmp, "biases.gz")
        write_biases(biasfile, *bias)
    model = ref.FitHiC(os.path.join(tmp, "lib"), resolution, n_bins=n_bins, max_dist=max_dist, min_dist=min_dist)
    lo, hi = model.min_dist, model.max_dist
    main = ref.generate_FragPairs(frags, resolution, lo, hi, False)
    bias_dic = ref.read_bias_file(biasfile, False) if biasfile != "none" else {}
    main = ref.read_interactions(main, inter, lo, hi, False)
    x, y, yerr = ref.calculate_probabilities(main, n_bins, resolution, lo, hi, os.path.join(tmp, "lib.fithic_pass1"), False)
    spline_x, new_y, residual = ref.fit_spline(main, x, y, yerr, inter, os.path.join(tmp, "lib.spline_pass1"),
                                               bias_dic, resolution, lo, hi, False)
    out = parse_significances(os.path.join(tmp, "lib.spline_pass1.res%d.significances.txt.gz" % resolution))
    keys = sorted(main)
    return {
        "min_dist": lo, "max_dist": hi,
        "possible": np.array([main[k][0] for k in keys], np.int64),
        "observed": np.array([main[k][1] for k in keys], np.int64),
        "x": np.array(x, np.float64), "y": np.array(y, np.float64),
        "spline_x": np.array(spline_x, np.int64), "spline_y": np.array(new_y, np.float64),
        "residual": float(residual),
        "S": int(ref.observedIntraInRangeSum),
        "intra_in_range_count": int(ref.observedIntraInRangeCount),
        "intra_all_sum": int(ref.observedIntraAllSum), "intra_all_count": int(ref.observedIntraAllCount),
        "inter_all_sum": int(ref.observedInterAllSum), "inter_all_count": int(ref.observedInterAllCount),
        "min_obs_dist": int(ref.minObservedGenomicDist), "max_obs_dist": int(ref.maxObservedGenomicDist),
        "max_possible_dist": int(ref.maxPossibleGenomicDist),
        "possible_intra_in_range": int(ref.possibleIntraInRangeCount),
        "possible_intra_all": int(ref.possibleIntraAllCount),
        "possible_inter_all": int(ref.possibleInterAllCount),
        "out": out,
        "ref_module": ref,
    }


def run_reference_two_pass(frag_chrom, frag_mid, chr1, mid1, chr2, mid2, count, resolution, keep_mask,
                           n_bins=100, min_dist=-1, max_dist=-1, bias=None):
    """Pass 2 composed from the reference's own functions: pass 1, drop the rows with
    p <= 1/possibleIntraInRangeCount from the interactions file, recompute the statistics on the filtered
    file in a FRESH module, then let fit_spline score the FULL file.  keep_mask: which input rows pass 1
    emitted (validated against the oracle) - needed to map output rows back to input rows."""
    import tempfile
    ref1 = run_reference_pass(frag_chrom, frag_mid, chr1, mid1, chr2, mid2, count, resolution, n_bins, min_dist, max_dist, bias)
    thr = 1.0 / ref1["possible_intra_in_range"]
    idx_kept = np.nonzero(keep_mask)[0]
    assert len(idx_kept) == len(ref1["out"]["p"])
    outlier_idx = idx_kept[ref1["out"]["p"] <= thr]
    sel = np.ones(len(count), dtype=bool)
    sel[outlier_idx] = False
    ref = ref_loader.load_reference_fithic()
    tmp = tempfile.mkdtemp(prefix="bbk_ref2_")
    full = os.path.join(tmp, "interactions_full.gz")
    filt = os.path.join(tmp, "interactions_filtered.gz")
    frags = os.path.join(tmp, "fragments.gz")
    write_interactions(full, chr1, mid1, chr2, mid2, count)
    write_interactions(filt, np.asarray(chr1)[sel], np.asarray(mid1)[sel], np.asarray(chr2)[sel], np.asarray(mid2)[sel], np.asarray(count)[sel])
    write_fragments(frags, frag_chrom, frag_mid)
    bias_dic = {}
    if bias is not None:
        bf = os.path.join(tmp, "biases.gz")
        write_biases(bf, *bias)
        bias_dic = ref.read_bias_file(bf, False)
    lo, hi = ref1["min_dist"], ref1["max_dist"]
    main = ref.generate_FragPairs(frags, resolution, lo, hi, False)
    main = ref.read_interactions(main, filt, lo, hi, False)
    x, y, yerr = ref.calculate_probabilities(main, n_bins, resolution, lo, hi, os.path.join(tmp, "lib.fithic_pass2"), False)
    spline_x, new_y, residual = ref.fit_spline(main, x, y, yerr, full, os.path.join(tmp, "lib.spline_pass2"), bias_dic,
                                               resolution, lo, hi, False)
    out = parse_significances(os.path.join(tmp, "lib.spline_pass2.res%d.significances.txt.gz" % resolution))
    keys = sorted(main)
    return {"threshold": thr, "n_outliers": int(len(outlier_idx)), "outlier_idx": outlier_idx,
            "observed": np.array([main[k][1] for k in keys], np.int64), "S": int(ref.observedIntraInRangeSum),
            "x": np.array(x), "y": np.array(y), "spline_x": np.array(spline_x, np.int64), "spline_y": np.array(new_y),
            "out": out}


def run_reference_decimate(map5, resolution):
    """FithicContactMap.decimate (datatypes.pyx:317-339) itself: the method's source is read from where it lies, exec'd as
    plain Python (the body uses nothing of Cython) on a stand-in object, with ONE edit - the two `/` of line 331 become
    `//`, Python 2's integer division on the integer midpoints.  Returns the (g, 5) map it leaves in self.map."""
    import textwrap
    import types
    path = os.path.join(ref_loader.REFERENCE_ROOT, "blueberry", "datatypes.pyx")
    lines = open(path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.strip().startswith("def decimate(self"))
    end = next(i for i in range(start + 1, len(lines)) if lines[i].strip().startswith("def "))
    body = lines[start:end]
    old = "/ resolution * resolution - resolution/2"
    hit = [i for i, l in enumerate(body) if old in l]
    if len(hit) != 1:
        raise RuntimeError("reference datatypes.pyx decimate changed; the edit no longer applies")
    body[hit[0]] = body[hit[0]].replace(old, "// resolution * resolution - resolution//2")
    src = textwrap.dedent("\n".join(body).replace("\t", "    "))
    ns = {"numpy": np}
    exec(compile(src, path, "exec"), ns)
    obj = types.SimpleNamespace(map=np.array(map5, dtype=np.float64, copy=True), resolution=None, regions=None)
    ns["decimate"](obj, resolution)
    return obj.map


def run_reference_contact_map(pos1, pos2, count, kr_norm, kr_expected, resolution):
    """The reference's ContactMap (datatypes.pyx:88-171), compiled verbatim: constructor (dense scatter of the RAWobserved
    rows) then normalize().  Its input paths come from module-level templates (datatypes.pyx:27-29); they are pointed at
    temporary files here.  Returns (matrix after __init__, matrix after normalize, regions, n_bins)."""
    dt = ref_loader.load_reference_datatypes()
    tmp = tempfile.mkdtemp(prefix="bbk_refcm_")
    dt.RAW_DIR = os.path.join(tmp, "{0}.chr{1}.{2}.RAWobserved")
    dt.KR_NORM = os.path.join(tmp, "{0}.chr{1}.{2}.KRnorm")
    dt.KR_EXP = os.path.join(tmp, "{0}.chr{1}.{2}.KRexpected")
    for tag in (resolution // 1000, resolution / 1000):            # Python 2 formats an int here, Python 3 a float
        for path, vec in ((dt.KR_NORM.format("cell", 1, tag), kr_norm), (dt.KR_EXP.format("cell", 1, tag), kr_expected)):
            with open(path, "w") as fh:
                for v in vec:
                    fh.write("%r\n" % float(v))
        with open(dt.RAW_DIR.format("cell", 1, tag), "w") as fh:
            for a, b, c in zip(pos1, pos2, count):
                fh.write("%r\t%r\t%r\n" % (float(a), float(b), float(c)))
    cm = dt.ContactMap("cell", 1, resolution)
    before = np.array(cm.matrix, copy=True)
    regions = np.array(cm.regions, copy=True)
    cm.normalize()
    return before, np.array(cm.matrix, copy=True), regions, int(cm.n_bins)


def write_reference_significances(map5, chromosome, resolution):
    """A significances file (fithic.py:410-435 format) with these rows where the compiled reference's FithicContactMap will look
    for it: its DATA_DIR template (a module global, datatypes.pyx:26) is pointed at a temporary directory."""
    import gzip
    dt = ref_loader.load_reference_datatypes()
    tmp = tempfile.mkdtemp(prefix="bbk_refex_")
    dt.DATA_DIR = os.path.join(tmp, "{0}.chr{1}.res{2}.significances.txt.gz")
    with gzip.open(dt.DATA_DIR.format("cell", chromosome, resolution), "wt") as fh:
        fh.write("chr1\tfragmentMid1\tchr2\tfragmentMid2\tcontactCount\tp-value\tq-value\n")
        for mid1, mid2, cnt, p, q in np.asarray(map5, dtype=np.float64):
            fh.write("chr{0}\t{1}\tchr{0}\t{2}\t{3}\t{4}\t{5}\n".format(chromosome, int(mid1), int(mid2), int(cnt), repr(float(p)), repr(float(q))))
    return dt


def reference_map_as_read(chromosome, resolution):
    """The table as the reference's FithicContactMap holds it after reading the file write_reference_significances wrote
    (datatypes.pyx:314: pandas' C parser with its default float conversion, which is NOT round-trip exact - the p-values it
    yields can differ from the written ones in the last bit, so goldens store THIS table as the input)."""
    dt = ref_loader.load_reference_datatypes()
    return np.array(dt.FithicContactMap("cell", chromosome, resolution).map, copy=True)


def run_reference_extract_contacts(chromosome, resolution, alpha=None, n_regions=None):
    """utils.extract_contacts (utils.py:31-90) itself, on the file write_reference_significances wrote: the function's source is
    read from where it lies and exec'd with a closed list of edits - its two Python-2 print statements (:61, :66) become calls,
    `e.message` (:66) becomes `e` - against the reference's own FithicContactMap (datatypes.pyx compiled verbatim) and its own
    count_band_regions (blueberry.pyx compiled verbatim)."""
    import textwrap
    dt = ref_loader.load_reference_datatypes()
    bb = ref_loader.load_reference_cython()
    path = os.path.join(ref_loader.REFERENCE_ROOT, "blueberry", "utils.py")
    lines = open(path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("def extract_contacts("))
    end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith("def "))
    body = lines[start:end]
    edits = 0
    for i, l in enumerate(body):
        st = l.strip()
        if st.startswith("print "):
            body[i] = l[:len(l) - len(l.lstrip())] + "print(" + st[len("print "):].replace("e.message", "e") + ")"
            edits += 1
    if edits != 2:
        raise RuntimeError("reference utils.py extract_contacts changed; the edit list no longer applies")
    src = textwrap.dedent("\n".join(body).replace("\t", "    "))
    ns = {"numpy": np, "FithicContactMap": dt.FithicContactMap, "count_band_regions": bb.count_band_regions,
          "HIGH_FITHIC_CUTOFF": 10000000, "LOW_FITHIC_CUTOFF": 25000}
    exec(compile(src, path, "exec"), ns)
    return ns["extract_contacts"]("cell", chromosome, resolution, alpha, n_regions)
