"""libbbkio.so (include/bbk_io.h): the significances writer against a line-by-line restatement of fithic.py:410-435.
Host code only - these tests need no GPU."""
import ctypes
import gzip
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def io():
    from blueberry_b200 import build, _io
    build.build_io()
    _io.load()
    return _io


def _reference_text(names, c1, m1, c2, m2, cnt, p, q):
    """What fithic.py:410-435 writes (the reference never has q: it writes -1; with q we format it like p)."""
    out = ["chr1\tfragmentMid1\tchr2\tfragmentMid2\tcontactCount\tp-value\tq-value\n"]
    for i in range(len(p)):
        if p[i] <= 1:
            out.append("{}\t{}\t{}\t{}\t{}\t{}\t{}\n".format(names[c1[i]], m1[i], names[c2[i]], m2[i], cnt[i], p[i],
                                                            -1 if q is None else q[i]))
    return "".join(out)


def test_header_symbols_are_exported(io):
    text = open(os.path.join(ROOT, "include", "bbk_io.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(bbkio_[a-z0-9_]+)\s*\(", text)))
    assert names == sorted(io.SIGNATURES)
    lib = ctypes.CDLL(io.LIB_PATH)
    for n in names:
        assert hasattr(lib, n)


def test_double_formatting_is_numpys(io):
    rng = np.random.default_rng(1)
    vals = [0.0, -0.0, 1.0, 0.5, 1e-4, 9.999e-5, 1e-5, 1e16, 9999999999999998.0, 1e15, 123456789.123, 5e-324,
            1.7976931348623157e308, 2.2250738585072014e-308, 0.1, 1 / 3, 1e22, 1e-300, 0.0001234, 1e-7, 3.0e-310, float("nan")]
    vals += list(np.exp(rng.uniform(-700, 700, 20000)))
    vals += list(rng.random(20000))
    bits = rng.integers(0, 2 ** 63 - 1, 20000, dtype=np.int64).view(np.float64)
    vals += [float(x) for x in bits if np.isfinite(x)]
    for x in vals:
        assert io.format_double(x) == "{}".format(np.float64(x)), repr(x)


@pytest.mark.parametrize("n,threads,with_q", [(0, 1, False), (1, 1, True), (5000, 1, False), (300001, 3, True), (300001, 0, False)])
def test_writer_matches_the_reference_loop(io, tmp_path, n, threads, with_q):
    rng = np.random.default_rng(n + threads)
    names = ["chr1", "chrX", "chr10_random"]
    c1 = rng.integers(0, 3, n).astype(np.int32)
    c2 = rng.integers(0, 3, n).astype(np.int32)
    m1 = rng.integers(0, 250_000_000, n)
    m2 = m1 + rng.integers(-5, 2000, n) * 5000
    cnt = rng.integers(0, 5000, n)
    p = np.exp(rng.uniform(-720, 0.0, n))
    p[rng.random(n) < 0.3] = 1.0
    p[rng.random(n) < 0.05] = np.nan               # rows the reference drops (fithic.py:434)
    p[rng.random(n) < 0.01] = 1.5
    p[rng.random(n) < 0.01] = 0.0
    q = np.minimum(p * 3.7, 1.0) if with_q else None
    path = str(tmp_path / "out.significances.txt.gz")
    rows = io.write_significances(path, names, c1, m1, c2, m2, cnt, p, q, threads=threads, level=1)
    want = _reference_text(names, c1, m1, c2, m2, cnt, p, q)
    with gzip.open(path, "rt") as fh:
        got = fh.read()
    assert got == want
    assert rows == int((p <= 1).sum())
    if n:
        import pandas as pd            # what FithicContactMap does with the file (datatypes.pyx:314)
        m = pd.read_csv(path, sep="\t", usecols=[1, 3, 4, 5, 6], engine="c", dtype="float64").values
        assert m.shape == (rows, 5)


def test_single_chromosome_shortcut_and_errors(io, tmp_path):
    p = np.array([0.5, np.nan, 1.0])
    path = str(tmp_path / "one.gz")
    assert io.write_significances(path, ["chr7"], None, [10, 20, 30], None, [15, 25, 35], [1, 2, 0], p) == 2
    assert gzip.open(path, "rt").read().splitlines()[1:] == ["chr7\t10\tchr7\t15\t1\t0.5\t-1", "chr7\t30\tchr7\t35\t0\t1.0\t-1"]
    with pytest.raises(io.BbkIoError):
        io.write_significances(str(tmp_path / "nodir" / "x.gz"), ["chr7"], None, [1], None, [2], [3], np.array([0.1]))
    with pytest.raises(io.BbkIoError):
        io.write_significances(path, ["chr7"], np.array([1], np.int32), [1], np.array([0], np.int32), [2], [3], np.array([0.1]))


def _ref_parse(path):
    """fithic.py:243-247, line by line."""
    rows = []
    with gzip.open(path, "rt") as fh:
        for line in fh:
            ch1, mid1, ch2, mid2, contactCount = line.rstrip().split()
            rows.append((ch1, int(mid1), ch2, int(mid2), int(contactCount)))
    return rows


@pytest.mark.parametrize("n,threads", [(0, 1), (3, 1), (400000, 3), (400000, 0)])
def test_reader_matches_the_reference_loop(io, tmp_path, n, threads):
    rng = np.random.default_rng(n + 7)
    names = np.array(["chr1", "chr10", "chrX", "scaffold_12"])
    a = names[rng.integers(0, 4, n)]
    b = names[rng.integers(0, 4, n)]
    m1 = rng.integers(-5, 250_000_000, n)
    m2 = rng.integers(0, 250_000_000, n)
    cnt = rng.integers(0, 100000, n)
    seps = ["\t", " ", "  \t "]
    path = str(tmp_path / "inter.gz")
    with gzip.open(path, "wt", compresslevel=1) as fh:          # rows of the shapes str.split() accepts
        chunks = []
        for i in range(n):
            s = seps[i % 3]
            chunks.append("%s%s%s%d%s%s%s%d%s%d%s\n" % ("  " if i % 5 == 0 else "", a[i], s, m1[i], s, b[i], s, m2[i], s, cnt[i],
                                                        " \r" if i % 7 == 0 else ""))
        text = "".join(chunks)
        fh.write(text[:-1] if n == 3 else text)                 # n == 3: last line without a newline
    got_names, c1, g1, c2, g2, gc = io.read_interactions(path, threads=threads)
    want = _ref_parse(path)
    assert len(want) == n == len(g1)
    if n:
        gn = np.array(got_names)
        assert [tuple(r) for r in zip(gn[c1], g1, gn[c2], g2, gc)] == want
        first = []
        for r in want:                                          # ids follow first appearance (chr1 before chr2 in a row)
            for nm in (r[0], r[2]):
                if nm not in first:
                    first.append(nm)
        assert got_names == first


def test_reader_accepts_plain_text_and_concatenated_gzip_members(io, tmp_path):
    rows = "chr1\t10\tchr1\t20\t3\nchr2 5 chr1 7 0\n"
    plain = tmp_path / "plain.txt"
    plain.write_text(rows)
    names, c1, m1, c2, m2, cnt = io.read_interactions(str(plain))
    assert names == ["chr1", "chr2"] and m1.tolist() == [10, 5] and cnt.tolist() == [3, 0] and c2.tolist() == [0, 0]
    multi = tmp_path / "multi.gz"
    with open(multi, "wb") as fh:
        fh.write(gzip.compress(rows.encode()))
        fh.write(gzip.compress(b"chrX\t1\tchrX\t2\t9\n"))
    names, c1, m1, c2, m2, cnt = io.read_interactions(str(multi))
    assert names == ["chr1", "chr2", "chrX"] and cnt.tolist() == [3, 0, 9]


@pytest.mark.parametrize("bad,msg", [("chr1\t1\tchr1\t2\n", "not enough values to unpack (expected 5, got 4)"),
                                     ("chr1\t1\tchr1\t2\t3\t4\n", "too many values to unpack (expected 5)"),
                                     ("chr1\t1.5\tchr1\t2\t3\n", "invalid literal for int() with base 10"),
                                     ("chr1\t1\tchr1\t2\t3\n\nchr1\t1\tchr1\t2\t3\n", "not enough values to unpack (expected 5, got 0)")])
def test_reader_fails_like_the_reference_on_malformed_rows(io, tmp_path, bad, msg):
    path = str(tmp_path / "bad.gz")
    with gzip.open(path, "wt") as fh:
        fh.write("chr1\t1\tchr1\t2\t3\n" * 10 + bad)
    with pytest.raises(ValueError) as e:
        io.read_interactions(path)
    assert msg in str(e.value)
    with pytest.raises(ValueError):                             # and so does the reference's loop
        _ref_parse(path)
    with pytest.raises(io.BbkIoError):
        io.read_interactions(str(tmp_path / "missing.gz"))


def test_unpack_scores_against_a_handmade_packing():
    """bbkio_unpack_scores (the host half of bbk_pack_scores): codes 0 / 1 / 2 / 3 = (1.0, 1.0) / (NaN, NaN) / (p, 1.0) / (p, q),
    values taken from the chunk's blocks in row order; a chunk table that disagrees with the codes is an error."""
    from blueberry_b200 import _io
    rng = np.random.default_rng(5)
    m = 4096 + 77                                             # two chunks, the second one partial
    codes = np.zeros((m + 15) // 16, dtype=np.uint32)
    exp_p, exp_q = np.ones(m), np.ones(m)
    lists = [([], []), ([], [])]
    for r in range(m):
        cd = int(rng.integers(0, 4))
        codes[r >> 4] |= np.uint32(cd << (2 * (r % 16)))
        vp, vq = lists[r // 4096]
        if cd == 1:
            exp_p[r] = exp_q[r] = np.nan
        elif cd >= 2:
            exp_p[r] = rng.random(); vp.append(exp_p[r])
            if cd == 3:
                exp_q[r] = rng.random(); vq.append(exp_q[r])
    chunks = np.zeros(2, dtype=[("bp", "<u8"), ("bq", "<u8"), ("np", "<u4"), ("nq", "<u4")])
    # the second chunk's values come FIRST in the lists (chunks land in no particular order)
    chunks["bp"] = [len(lists[1][0]), 0]; chunks["bq"] = [len(lists[1][1]), 0]
    chunks["np"] = [len(lists[0][0]), len(lists[1][0])]; chunks["nq"] = [len(lists[0][1]), len(lists[1][1])]
    vals_p, vals_q = np.array(lists[1][0] + lists[0][0]), np.array(lists[1][1] + lists[0][1])
    p, q = _io.unpack_scores(codes, chunks.view(np.uint8), vals_p, vals_q, m, threads=2)
    assert np.array_equal(p.view(np.uint64), exp_p.view(np.uint64)) and np.array_equal(q.view(np.uint64), exp_q.view(np.uint64))
    p, q = _io.unpack_scores(codes, chunks.view(np.uint8), vals_p, vals_q, m, want_q=False)
    assert np.array_equal(p.view(np.uint64), exp_p.view(np.uint64)) and q is None
    chunks["np"][0] += 1
    with pytest.raises(_io.BbkIoError):
        _io.unpack_scores(codes, chunks.view(np.uint8), vals_p, vals_q, m)

