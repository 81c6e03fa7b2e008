"""Host-side logic of the drop-in (no GPU): fragment bookkeeping, bias tables, sharding, interface shape."""
import inspect

import numpy as np
import pytest

from helpers import PASS_CASES, load_golden


@pytest.mark.parametrize("name", PASS_CASES)
def test_fragment_bookkeeping_matches_reference(name):
    from blueberry_b200.fithic import _frag_info
    g = load_golden(name)
    info = _frag_info(g["frag_chrom"], g["frag_mid"], int(g["resolution"]))
    assert info.max_possible == int(g["ref_max_possible_dist"])
    assert info.nkeys == len(g["ref_possible"])
    assert info.possible_intra_all == int(g["ref_possible_intra_all"])
    assert info.possible_inter_all == int(g["ref_possible_inter_all"])
    # possible[k] = sum_c (k*R <= max_frag_c ? n_c - k : 0) is what bbk_possible_pairs computes on the device
    R = int(g["resolution"])
    k = np.arange(info.nkeys)
    poss = np.zeros(info.nkeys, dtype=np.int64)
    for n, mf in zip(info.n_frags, info.max_frag):
        poss += np.where(k * R <= mf, n - k, 0)
    assert np.array_equal(poss, g["ref_possible"])


def test_bias_dict_from_arrays_follows_read_bias_file():
    from blueberry_b200.fithic import _bias_dict_from_arrays
    from oracle import fithic_oracle as fo
    g = load_golden("pass_messy")
    mine = _bias_dict_from_arrays(g["bias_chrom"], g["bias_mid"], g["bias_val"])
    ref, _ = fo.read_bias_arrays(g["bias_chrom"], g["bias_mid"], g["bias_val"])
    mine = {c: dict(zip(mids.tolist(), vals.tolist())) for c, (mids, vals) in mine.items()}      # (mids, values) arrays per chromosome
    assert mine == {c: {m: float(v) for m, v in sub.items()} for c, sub in ref.items()}


def test_interface_mirrors_the_reference():
    from blueberry_b200 import fithic as f
    assert list(inspect.signature(f.FitHiC.__init__).parameters) == ["self", "libname", "resolution", "n_bins", "n_passes", "max_dist", "min_dist"]
    d = inspect.signature(f.FitHiC.__init__).parameters
    assert (d["n_bins"].default, d["n_passes"].default, d["max_dist"].default, d["min_dist"].default) == (100, 2, -1, -1)
    assert list(inspect.signature(f.FitHiC.fit_transform).parameters) == ["self", "interactions", "fragments", "biases", "verbose"]
    assert list(inspect.signature(f.fithic).parameters)[:10] == ["libname", "resolution", "n_bins", "min_dist", "max_dist", "n_passes",
                                                                  "interactions", "frags", "biases", "verbose"]
    assert list(inspect.signature(f.benjamini_hochberg_correction).parameters) == ["p_values", "num_total_tests"]
    m = f.FitHiC("x", 5000)
    assert (m.max_dist, m.min_dist) == (10000000, 0)              # fithic.py:82-83
    assert f.in_range_check(5, 0, 10) and not f.in_range_check(0, 0, 10) and f.in_range_check(10, 0, 10)
    assert f.in_range_check(10 ** 12, -1, -1)
    from blueberry_b200 import blueberry as b
    assert list(inspect.signature(b.benjamini_hochberg).parameters) == ["p_values", "n"]
    from blueberry_b200 import utils
    assert (utils.LOW_FITHIC_CUTOFF, utils.HIGH_FITHIC_CUTOFF) == (25000, 10000000)


def test_lpt_sharding_and_bands():
    from blueberry_b200 import sharding, synth
    R, K = 5000, 2000
    loads = [synth.n_pairs_of(synth.n_bins_of(L, R), K) for L in synth.HG19_LENGTHS]
    assert sum(loads) == 1169126271                               # BASELINE config 3
    for world in (2, 4, 8):
        owner = sharding.lpt_assign(loads, world)
        tot = [sum(l for l, o in zip(loads, owner) if o == r) for r in range(world)]
        assert max(tot) / (sum(tot) / world) < 1.06
    nb = synth.n_bins_of(synth.HG19_LENGTHS[0], 1000)
    bands = sharding.split_bands(nb, 2000, 8)
    assert bands[0][0] == 0 and bands[-1][1] == nb and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
    sizes = []
    for lo, hi in bands:
        rows = np.arange(lo, hi)
        sizes.append(int(np.minimum(2000, nb - 1 - rows).sum() + len(rows)))
    assert sum(sizes) == synth.n_pairs_of(nb, 2000) == 496750251   # BASELINE config 4
    assert max(sizes) / min(sizes) < 1.001


def test_plan_shards_covers_every_record_once_and_balances():
    """distributed.plan_shards on the hg19 genome at 5 kb (BASELINE config 3's shape): every record of every chromosome
    lands on exactly one rank, pieces are contiguous inside a chromosome, loads are equal to 4 records, and at most
    world-1 chromosomes are split."""
    from blueberry_b200 import synth
    from blueberry_b200.distributed import layout_rows, plan_shards, shard_rows
    K = 2000
    pairs = [synth.n_pairs_of(synth.n_bins_of(L, 5000), K) for L in synth.HG19_LENGTHS]
    assert sum(pairs) == 1169126271                                     # cfg3
    for world in (1, 2, 4, 8):
        plan = plan_shards(pairs, world)
        loads = [sum(n for _, _, n in r) for r in plan]
        assert sum(loads) == sum(pairs) and max(loads) - min(loads) <= 8
        split = 0
        for c, n in enumerate(pairs):
            segs = sorted((f, k) for r in plan for (cc, f, k) in r if cc == c)
            pos = 0
            for f, k in segs:
                assert f == pos
                pos += k
            assert pos == n
            split += len(segs) > 1
        assert split <= world - 1
    lpt = plan_shards(pairs, 8, mode="lpt")
    assert sorted(c for r in lpt for (c, _, _) in r) == list(range(23)) and all(f == 0 for r in lpt for (_, f, _) in r)
    # arbitrary tables: contiguous row slices, 4-aligned cuts
    for n in (0, 1, 7, 1000, 1003):
        got = [shard_rows(n, 3, r) for r in range(3)]
        assert got[0][0] == 0 and got[-1][1] == n and all(a[1] == b[0] for a, b in zip(got[:-1], got[1:]))
        assert all(lo % 4 == 0 for lo, _ in got if lo < n)
    assert layout_rows([5, 0, 8, 3]) == ([0, 8, 8, 16], 20)
