"""The numpy oracle must reproduce what the reference itself produced (tests/golden/*.npz)."""
import numpy as np
import pytest

from oracle import fithic_oracle as fo
from helpers import PASS_CASES, load_golden, golden_bias_dict


@pytest.mark.parametrize("name", PASS_CASES)
def test_pass_matches_reference_output(name):
    g = load_golden(name)
    R = int(g["resolution"])
    lo, hi = int(g["ref_min_dist"]), int(g["ref_max_dist"])
    o = fo.fithic_arrays(g["frag_chrom"], g["frag_mid"], g["chr1"], g["mid1"], g["chr2"], g["mid2"], g["count"],
                         R, int(g["n_bins"]), lo, hi, bias=golden_bias_dict(g))
    # integers: bit exact
    assert np.array_equal(o.frag.possible, g["ref_possible"])
    assert np.array_equal(o.contacts.observed, g["ref_observed"])
    for k in ("S", "intra_in_range_count", "intra_all_sum", "intra_all_count", "inter_all_sum",
              "inter_all_count", "min_obs_dist", "max_obs_dist"):
        assert getattr(o.contacts, k) == int(g["ref_" + k]), k
    assert o.frag.max_possible_dist == int(g["ref_max_possible_dist"])
    assert o.frag.possible_intra_in_range == int(g["ref_possible_intra_in_range"])
    assert o.frag.possible_intra_all == int(g["ref_possible_intra_all"])
    assert o.frag.possible_inter_all == int(g["ref_possible_inter_all"])
    # floats: the oracle makes the same library calls on the same values -> bit exact too
    assert np.array_equal(np.array(o.x), g["ref_x"])
    assert np.array_equal(np.array(o.y), g["ref_y"])
    assert o.k0 * R == int(g["ref_spline_x"][0]) and len(o.spline_y) == len(g["ref_spline_x"])
    assert np.array_equal(o.spline_y, g["ref_spline_y"])
    assert o.residual == float(g["ref_residual"])
    # the rows the reference wrote, in order, with its p-values (q column is the literal -1, fithic.py:435)
    keep = o.keep
    assert np.array_equal(g["mid1"][keep], g["ref_out_mid1"])
    assert np.array_equal(g["mid2"][keep], g["ref_out_mid2"])
    assert np.array_equal(g["count"][keep], g["ref_out_count"])
    assert np.array_equal(o.p[keep], g["ref_out_p"])
    assert (g["ref_out_q"] == -1).all()


def test_spline_index_closed_form_is_bisect():
    g = load_golden("pass_messy")
    R = int(g["resolution"])
    k0, L = int(g["ref_spline_x"][0]) // R, len(g["ref_spline_x"])
    d = (g["mid2"].astype(np.int64) - g["mid1"])[:5000]
    a = fo.spline_index(d, k0, L, R, float(g["ref_x"].min()), float(g["ref_x"].max()))
    b = fo.spline_index_closed_form(d, k0, L, R)
    assert np.array_equal(a, b)


def test_bh_and_band_match_reference():
    g = load_golden("bh_band")
    for tag in ("a", "b", "c", "doc"):
        p, n = g["p_" + tag], int(g["n_" + tag])
        assert np.array_equal(fo.benjamini_hochberg_correction(p, n), g["q_py_" + tag])
        assert np.array_equal(fo.benjamini_hochberg_sorted(np.sort(p), n), g["q_cy_sorted_" + tag])
    assert fo.count_band_regions(g["regions_sorted"]) == int(g["band_sorted"])
    assert fo.count_band_regions(g["regions_shuffled"]) == int(g["band_shuffled"])


def test_bh_is_forward_running_max_not_textbook():
    # the reference's BH differs from the textbook reverse cumulative-min
    rng = np.random.default_rng(3)
    p = rng.random(1000)
    q = fo.benjamini_hochberg_correction(p, 1000)
    order = np.argsort(p)
    textbook = np.minimum.accumulate((np.sort(p) * 1000 / np.arange(1, 1001))[::-1])[::-1]
    tb = np.empty(1000); tb[order] = np.minimum(textbook, 1)
    assert (q != tb).sum() > 100


def test_two_pass_oracle_matches_composed_reference():
    """Pass 2 has no reference code; the oracle follows the composition of reference functions of SURVEY 8c,
    and must reproduce what those functions produced when run for real (tests/golden/pass2_bias_dense.npz)."""
    g = load_golden("pass_bias_dense")
    g2 = load_golden("pass2_bias_dense")
    R = int(g["resolution"])
    r1, r2, outlier, thr = fo.fithic_two_pass_arrays(g["frag_chrom"], g["frag_mid"], g["chr1"], g["mid1"], g["chr2"], g["mid2"],
                                                     g["count"], R, int(g["n_bins"]), int(g["ref_min_dist"]), int(g["ref_max_dist"]),
                                                     bias=golden_bias_dict(g))
    assert thr == float(g2["threshold"]) and int(outlier.sum()) == int(g2["n_outliers"])
    assert np.array_equal(np.nonzero(outlier)[0], g2["outlier_idx"])
    assert np.array_equal(r2.contacts.observed, g2["ref2_observed"]) and r2.contacts.S == int(g2["ref2_S"])
    assert np.array_equal(np.array(r2.x), g2["ref2_x"]) and np.array_equal(np.array(r2.y), g2["ref2_y"])
    assert np.array_equal(r2.spline_y, g2["ref2_spline_y"])
    assert np.array_equal(g["mid1"][r2.keep], g2["ref2_out_mid1"])
    assert np.array_equal(r2.p[r2.keep], g2["ref2_out_p"])


def test_decimate_oracle_matches_reference_method_output():
    """oracle/datatypes_oracle.decimate against tests/golden/decimate.npz (FithicContactMap.decimate run from the
    reference's own method source, datatypes.pyx:317-339, by oracle/make_golden.py)."""
    from oracle import datatypes_oracle as do
    g = load_golden("decimate")
    out5 = do.decimate(g["map_in"], 5000)
    assert np.array_equal(out5, g["ref_5000"])
    assert np.array_equal(do.decimate(out5, 25000), g["ref_25000_of_5000"])
    # the rounding rule on a few hand-checked values: (mid + r) // r * r - r // 2
    m = np.array([[500.0, 4500.0, 1, 0.5, 1.0], [4999.0, 5000.0, 2, 0.5, 0.2], [0.0, 9999.0, 3, 0.25, 0.7]])
    d = do.decimate(m, 5000)
    assert d.tolist() == [[2500.0, 2500.0, 1.0, 0.5, 1.0], [2500.0, 7500.0, 5.0, 0.125, 0.2]]
    assert do.contacts(np.array([[1.0, 2.0, 3, 0.1, 0.01], [3.0, 4.0, 3, 0.1, 0.011]])).tolist() == [[1.0, 2.0]]


def test_contact_map_oracle_matches_compiled_reference_output():
    """oracle/datatypes_oracle.contact_map_dense / normalize_dense against tests/golden/contact_map.npz (the reference's
    ContactMap compiled verbatim and run by oracle/make_golden.py, datatypes.pyx:88-171)."""
    from oracle import datatypes_oracle as do
    g = load_golden("contact_map")
    m, regions = do.contact_map_dense(g["pos1"], g["pos2"], g["count"], int(g["ref_n_bins"]), int(g["resolution"]))
    assert np.array_equal(m, g["ref_matrix"])
    assert np.array_equal(regions, g["ref_regions"])
    assert np.array_equal(do.normalize_dense(m, g["kr_norm"], g["kr_expected"], int(g["ref_n_bins"])), g["ref_normalized"])
    kr = g["kr_norm"].copy()
    kr[5] = 0.0
    with pytest.raises(ZeroDivisionError):
        do.normalize_dense(m, kr, g["kr_expected"], int(g["ref_n_bins"]))


def test_extract_contacts_and_genome_qvalues_oracle_match_reference_output():
    """utils.extract_contacts (utils.py:31-90) and the genome-wide q-value composition around it: golden
    minted by executing the reference's own function source against its compiled FithicContactMap / count_band_regions /
    benjamini_hochberg (oracle/make_golden.py: _extract_case)."""
    import os
    from oracle import datatypes_oracle as do
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "extract_contacts.npz"))
    alpha = float(g["alpha"])
    maps = {int(c): g["map_%d" % c] for c in g["chroms"]}
    for c, m in maps.items():
        assert np.array_equal(do.extract_contacts(m, c, alpha), g["ref_contacts_%d" % c])
        assert np.array_equal(do.extract_contacts(m, c), g["ref_contacts_noalpha_%d" % c])
        assert do.count_band_regions(do.regions(m)) == int(g["ref_band_%d" % c])
    contacts, q, n = do.genome_qvalues(maps, alpha)
    assert n == int(g["ref_n"]) and np.array_equal(q, g["ref_q"])
    assert np.array_equal(contacts, np.concatenate([g["ref_contacts_%d" % c] for c in g["chroms"]]))

