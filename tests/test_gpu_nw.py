"""GPU parity: Needleman-Wunsch identity kernels (through the C ABI) against the oracle, bit-exact
(matches, alignment_length) per pair and identical doubles."""
import os

import numpy as np
import pytest

import dynaalign_b200 as da
from conftest import ALPHABET24, GOLDEN, TABLES, fingerprint, random_seqs, same_matrix
from oracle import port

pytestmark = pytest.mark.gpu


def check_stats(seqs, name="BLOSUM62", go=10, ge=4):
    gm, gl = da.nw_pair_stats(seqs, name, go, ge)
    wm, wl = port.nw_pair_stats(seqs, name, go, ge)
    bad = np.nonzero((gm != wm) | (gl != wl))[0]
    assert bad.size == 0, "first mismatch at packed pair %d: got (%d,%d) want (%d,%d)" % (
        bad[0], gm[bad[0]], gl[bad[0]], wm[bad[0]], wl[bad[0]])


def test_known_answers(golden):
    g = golden["nw_pep4"]
    assert same_matrix(da.similarityNW(g["sequences"]), np.array(g["matrix"]))
    o = golden["nw_order"]
    m = da.similarityNW([o["a"], o["b"]])
    assert m[0, 1] == o["ab"] and m[1, 0] == o["ab"]
    m = da.similarityNW([o["b"], o["a"]])
    assert m[0, 1] == o["ba"]  # lower index on rows: order-sensitive like the reference


def test_empty_strings_and_nan():
    seqs = ["", "AA", "", "ARND", "A"]
    m = da.similarityNW(seqs)
    assert same_matrix(m, port.similarityNW(seqs))
    assert np.isnan(m[0, 0]) and np.isnan(m[0, 2]) and m[0, 1] == 0.0 and m[1, 1] == 1.0
    assert np.isnan(da.similarityNW([""])[0, 0])


@pytest.mark.parametrize("table", TABLES)
def test_thread_kernel_short_rows(table):
    # rows of 1..32 residues take the thread-per-pair kernel (all strip heights 4..32)
    import zlib
    rng = np.random.default_rng(zlib.crc32(table.encode()) % 1000)  # (hash() of a str changes from run to run)
    seqs = random_seqs(rng, 70, 1, 32) + random_seqs(rng, 10, 33, 90) + ["", "A"]
    rng.shuffle(seqs)
    go, ge = int(rng.integers(0, 13)), int(rng.integers(0, 6))
    check_stats(seqs, table, go, ge)


@pytest.mark.parametrize("go,ge", [(0, 0), (0, 4), (0, 1), (1, 0), (10, 1), (10, 0), (3, 2), (40, 0)])
def test_short_probes_gap_penalty_corners(go, ge):
    # the two-rows thread kernel treats its first row specially (no vertical gap from the border row) and, like every
    # two-rows kernel, runs in the unsigned domain also when score + 2*gapExt is negative: zero and tiny penalties
    rng = np.random.default_rng(go * 31 + ge)
    seqs = random_seqs(rng, 90, 1, 32, "ARNDCQEGHILKMFPSTWYV") + random_seqs(rng, 20, 1, 32) + random_seqs(rng, 6, 33, 70)
    seqs += ["W" * 16, "W" * 16, "C" * 32, "A", "AA", ""]
    rng.shuffle(seqs)
    check_stats(seqs, "BLOSUM62", go, ge)
    check_stats(seqs[:60], "BLOSUM100", go, ge)


@pytest.mark.parametrize("lo,hi", [(33, 64), (65, 130), (131, 260), (261, 400), (401, 600), (601, 768)])
def test_warp_kernel_all_strip_heights(lo, hi):
    rng = np.random.default_rng(lo)
    seqs = random_seqs(rng, 14, lo, hi, "ARNDCQEGHILKMFPSTWYV") + random_seqs(rng, 4, 1, 40)
    check_stats(seqs)


@pytest.mark.parametrize("R", [7, 9, 10, 12])
def test_cooperative_kernel_rows_385_to_768(R):
    # rows of 385..768 residues: two warps share one 64-lane wavefront (nw_warp2co_kernel), units of 128 columns;
    # lengths sit at the strip boundaries 64*(R-1)+1 .. 64*R, columns mix short, equal and > 1024 (fallback) lengths
    rng = np.random.default_rng(100 + R)
    lo, hi = max(385, 64 * (R - 1) + 1), 64 * R
    fam = "".join(random_seqs(rng, 1, hi, hi, "ARNDCQEGHILKMFPSTWYV"))
    seqs = [fam[: int(L)] for L in rng.integers(lo, hi + 1, size=70)]                    # related: long diagonal runs
    seqs += random_seqs(rng, 50, lo, hi, "ARNDCQEGHILKMFPSTWYV")                         # unrelated: gaps and ties
    seqs += [fam[:lo], fam[:hi], fam[5:hi], ""] + random_seqs(rng, 6, 1, 60) + random_seqs(rng, 2, 1025, 1040)
    rng.shuffle(seqs)
    check_stats(seqs)
    check_stats(seqs[:36], "BLOSUM62", 10, 1)  # negative slanted scores in the unsigned domain (score2_word)


@pytest.mark.parametrize("lo,hi", [(33, 96), (97, 200), (201, 330), (300, 384), (33, 384)])
def test_two_rows_kernel(lo, hi):
    # rows i, i+1 in the two 16-bit halves against the same column sequence (nw_rows2_kernel; by default only for inputs
    # large enough to fill the GPU with its units, forced here): adjacent rows of different and equal lengths, rows that
    # do not pair (too different, too short, too long), odd row counts, columns of every kind
    rng = np.random.default_rng(lo * 7 + hi)
    fam = "".join(random_seqs(rng, 1, hi, hi, "ARNDCQEGHILKMFPSTWYV"))
    seqs = [fam[: int(L)] for L in rng.integers(lo, hi + 1, size=40)] + random_seqs(rng, 41, lo, hi, "ARNDCQEGHILKMFPSTWYV")
    seqs += random_seqs(rng, 6, 1, 32) + ["", fam[:hi], fam[:hi], fam[:lo]] + random_seqs(rng, 3, 385, 700)
    rng.shuffle(seqs)
    os.environ["DYNA_NW_ROWS2"] = "1"
    try:
        check_stats(seqs)
        check_stats(seqs[:-1], "BLOSUM45", 3, 1)
    finally:
        del os.environ["DYNA_NW_ROWS2"]


@pytest.mark.parametrize("lo,hi", [(300, 384), (450, 640), (660, 768), (8, 32)])
def test_two_rows_kernels_long_columns_and_mixed_lengths(lo, hi):
    # the two-rows kernels stage column sequences of up to 2048 residues (1024 for the cooperative form from R = 11 on); a
    # longer one sends only ITS unit of the row pair to the single-row kernels.  Rows pair with the next compatible row
    # inside a window, not only with their neighbour: rows of other families are interleaved here.
    rng = np.random.default_rng(lo + hi)
    seqs = random_seqs(rng, 36, lo, hi, "ARNDCQEGHILKMFPSTWYV")
    seqs += [random_seqs(rng, 1, L, L, "ARNDCQEGHILKMFPSTWYV")[0] for L in (1025, 1500, 2048, 2049, 2600)]
    seqs += random_seqs(rng, 10, 1, 30) + random_seqs(rng, 8, 100, 200) + random_seqs(rng, 6, 400, 900) + [""]
    rng.shuffle(seqs)
    check_stats(seqs)
    os.environ["DYNA_NW_PAIR_WINDOW"] = "1"  # neighbours only: same results
    try:
        check_stats(seqs[:40])
    finally:
        del os.environ["DYNA_NW_PAIR_WINDOW"]


@pytest.mark.parametrize("lo,hi", [(769, 1000), (1000, 1152), (1153, 1536), (2000, 2304)])
def test_two_rows_multipass_kernel(lo, hi):
    # row pairs beyond the cooperative form (769..3072 residues): passes of 32*R rows, lane 31's bottom row through a
    # per-warp scratch line (nw_rows2mp_kernel).  Related sequences (long diagonal runs across the pass boundaries),
    # unrelated ones, rows that end in different passes (no pair), short and long columns, columns beyond the staging buffer.
    rng = np.random.default_rng(lo + 3 * hi)
    fam = "".join(random_seqs(rng, 1, hi, hi, "ARNDCQEGHILKMFPSTWYV"))
    seqs = [fam[: int(L)] for L in rng.integers(lo, hi + 1, size=14)] + random_seqs(rng, 10, lo, hi, "ARNDCQEGHILKMFPSTWYV")
    seqs += [fam[:hi], fam[3:hi], fam[:lo]] + random_seqs(rng, 8, 1, 400) + random_seqs(rng, 2, 2049, 2100) + [""]
    rng.shuffle(seqs)
    os.environ["DYNA_NW_ROWS2MP"] = "1"  # (by default only when these rows hold a tenth of the plan's cells: they do here)
    try:
        check_stats(seqs)
        a = da.nw_pair_stats(seqs[:20], "BLOSUM80", 7, 1)
        os.environ["DYNA_NW_ROWS2MP"] = "0"
        b = da.nw_pair_stats(seqs[:20], "BLOSUM80", 7, 1)
    finally:
        del os.environ["DYNA_NW_ROWS2MP"]
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    check_stats(seqs[:24])


def test_cooperative_kernel_equals_single_warp_kernels():
    # same input through the cooperative kernel and (DYNA_NW_CO=0) the tall-strip / multi-pass kernels
    rng = np.random.default_rng(77)
    seqs = random_seqs(rng, 150, 500, 700, "ARNDCQEGHILKMFPSTWYV")
    a = da.nw_pair_stats(seqs, "BLOSUM80", 7, 2)
    os.environ["DYNA_NW_CO"] = "0"
    try:
        b = da.nw_pair_stats(seqs, "BLOSUM80", 7, 2)
    finally:
        del os.environ["DYNA_NW_CO"]
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


def test_warp_kernel_exact_boundaries():
    # row lengths at the strip boundaries 32*R and 32*R+1, against short and long columns
    rng = np.random.default_rng(9)
    lens = [32, 33, 64, 65, 96, 97, 352, 353, 384, 385, 767, 768]
    seqs = [random_seqs(rng, 1, L, L, "ARNDCQEGHILKMFPSTWYV")[0] for L in lens] + random_seqs(rng, 3, 1, 5)
    check_stats(seqs, "BLOSUM50", 12, 3)


def test_multipass_long_rows():
    rng = np.random.default_rng(10)
    seqs = random_seqs(rng, 5, 769, 2100, "ARNDCQEGHILKMFPSTWYV") + random_seqs(rng, 4, 1, 300)
    rng.shuffle(seqs)
    check_stats(seqs)


def test_similar_sequences_exercise_gaps():
    # near-identical proteins with indels: long diagonal runs, ties and gap extensions
    rng = np.random.default_rng(12)
    base = random_seqs(rng, 1, 330, 330, "ARNDCQEGHILKMFPSTWYV")[0]
    seqs = [base]
    for _ in range(20):
        s = list(base)
        for _ in range(int(rng.integers(1, 12))):
            p = int(rng.integers(0, len(s)))
            r = rng.random()
            if r < 0.4:
                del s[p:p + int(rng.integers(1, 6))]
            elif r < 0.8:
                s[p:p] = list(random_seqs(rng, 1, 1, 5, "ARNDCQEGHILKMFPSTWYV")[0])
            else:
                s[p] = "W"
        seqs.append("".join(s))
    for name, go, ge in [("BLOSUM62", 10, 4), ("BLOSUM62", 0, 0), ("BLOSUM45", 1, 1), ("BLOSUM100", 12, 0), ("BLOSUM80", 3, 5)]:
        check_stats(seqs, name, go, ge)


def test_unslanted_recurrence_for_large_gap_extension():
    # gapExt too large for the slanted int8 profile -> the explicit "- ge" kernels
    rng = np.random.default_rng(13)
    seqs = random_seqs(rng, 12, 1, 120, "ARNDCQEGHILKMFPSTWYV")
    check_stats(seqs, "BLOSUM62", 30, 70)
    check_stats(seqs, "BLOSUM62", 5, 200)


def test_slanted_and_unslanted_agree(monkeypatch):
    rng = np.random.default_rng(14)
    seqs = random_seqs(rng, 20, 1, 200)
    a = da.nw_pair_stats(seqs, "BLOSUM62", 10, 4)
    monkeypatch.setenv("DYNA_NW_SLANT", "0")
    b = da.nw_pair_stats(seqs, "BLOSUM62", 10, 4)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


def test_row_range_slabs():
    rng = np.random.default_rng(15)
    seqs = random_seqs(rng, 60, 0, 90)
    wm, wl = port.nw_pair_stats(seqs)
    lens = [len(s) for s in seqs]
    bounds = da.partition_rows(len(seqs), 4, weights=lens, include_diagonal=True)
    gm = np.concatenate([da.nw_pair_stats(seqs, row_begin=int(bounds[s]), row_end=int(bounds[s + 1]))[0] for s in range(4)])
    gl = np.concatenate([da.nw_pair_stats(seqs, row_begin=int(bounds[s]), row_end=int(bounds[s + 1]))[1] for s in range(4)])
    assert (gm == wm).all() and (gl == wl).all()


def test_evp_full_matrix_fingerprint(golden, evp):
    m = da.similarityNW(evp)
    assert fingerprint(m) == golden["nw_evp_blosum62_10_4"]["fnv1a64"]
    for key, fp in golden.items():
        if key.startswith("nw_evp40_"):
            _, _, nm, go, ge = key.split("_")
            assert fingerprint(da.similarityNW(evp[:40], nm, int(go), int(ge))) == fp["fnv1a64"], key


def test_h3n2_first24(golden, h3n2):
    m = da.similarityNW(h3n2[:24])
    assert fingerprint(m) == golden["nw_h3n2_24"]["fnv1a64"]
    assert same_matrix(m, np.array(golden["nw_h3n2_24"]["matrix"]))


def test_h3n2_config2_full(golden, h3n2):
    # BASELINE config 2: similarityNW(h3n2sample$sequence[1:1000]); golden (matches, length) for all 500,500 pairs
    g = np.load(os.path.join(GOLDEN, "nw_h3n2_1000_stats.npz"))
    gm, gl = da.nw_pair_stats(h3n2)
    assert (gm == g["matches"]).all() and (gl == g["length"]).all()
    m = da.similarityNW(h3n2)
    assert (m == m.T).all() and (np.diag(m) == 1.0).all()
    for s in golden["nw_h3n2_1000_sample"]:
        assert m[s["i"], s["j"]] == s["sim"]


def test_packed16_and_32bit_kernels_agree(monkeypatch):
    # rows of 33..640 residues take the two-pairs-per-warp 16-bit kernel by default; DYNA_NW_PACK16=0 forces the 32-bit one
    rng = np.random.default_rng(16)
    seqs = random_seqs(rng, 25, 33, 640, "ARNDCQEGHILKMFPSTWYV") + random_seqs(rng, 6, 0, 40) + ["A" * 500, "W" * 640, "C" * 333]
    rng.shuffle(seqs)
    for name, go, ge in [("BLOSUM62", 10, 4), ("BLOSUM100", 0, 0), ("BLOSUM45", 12, 1), ("BLOSUM80", 2, 7)]:
        a = da.nw_pair_stats(seqs, name, go, ge)
        monkeypatch.setenv("DYNA_NW_PACK16", "0")
        b = da.nw_pair_stats(seqs, name, go, ge)
        monkeypatch.delenv("DYNA_NW_PACK16")
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    check_stats(seqs[:14], "BLOSUM62", 10, 4)


def test_packed16_range_guard():
    # long, self-similar sequences with a high-scoring table push the slanted score past int16: must fall back, stay exact
    seqs = ["W" * 640, "W" * 600, "W" * 64 + "A" * 500, "C" * 640]
    check_stats(seqs, "BLOSUM100", 1, 30)
    check_stats(seqs, "BLOSUM62", 10, 4)


def test_short_rows_against_longer_columns_many_shapes():
    # regression: the two-pairs-per-thread kernel once took 'U' for 'L' in the first DP row (compiler commuted a
    # VIMNMX.S16x2 with an immediate operand); single row sequence vs one or two column sequences of any length
    rng = np.random.default_rng(5)
    for trial in range(250):
        m, n = int(rng.integers(1, 33)), int(rng.integers(0, 60))
        seqs = random_seqs(rng, 1, m, m) + random_seqs(rng, 1 + trial % 2, n, n)
        go, ge = int(rng.integers(0, 13)), int(rng.integers(0, 6))
        name = TABLES[trial % 6]
        gm, gl = da.nw_pair_stats(seqs, name, go, ge)
        wm, wl = port.nw_pair_stats(seqs, name, go, ge)
        assert (gm == wm).all() and (gl == wl).all(), (seqs, name, go, ge)


def test_packed_multipass_rows(monkeypatch):
    # rows of 385..1500 residues: the packed kernel in several passes of 32*R rows (boundary rows through scratch)
    rng = np.random.default_rng(17)
    seqs = [random_seqs(rng, 1, L, L, "ARNDCQEGHILKMFPSTWYV")[0] for L in (385, 386, 640, 641, 767, 769, 1000, 1153, 1500)]
    seqs += random_seqs(rng, 6, 1, 700, "ARNDCQEGHILKMFPSTWYV") + ["", "W" * 900]
    rng.shuffle(seqs)
    check_stats(seqs)
    a = da.nw_pair_stats(seqs, "BLOSUM45", 3, 1)
    monkeypatch.setenv("DYNA_NW_PACK16", "0")
    b = da.nw_pair_stats(seqs, "BLOSUM45", 3, 1)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


def test_mixed_lengths_choose_kernels_per_unit():
    # a few very long sequences among short ones: units without them stay on the packed kernels, units with them
    # fall back to the 32-bit kernels -- results must not depend on which kernel a pair landed on
    rng = np.random.default_rng(18)
    seqs = random_seqs(rng, 150, 20, 400, "ARNDCQEGHILKMFPSTWYV")
    for pos, L in [(3, 1500), (77, 2300), (140, 1100)]:
        seqs[pos] = random_seqs(rng, 1, L, L, "ARNDCQEGHILKMFPSTWYV")[0]
    check_stats(seqs)


def test_columns_between_1024_and_2048_use_the_multipass_packed_kernel():
    rng = np.random.default_rng(19)
    seqs = random_seqs(rng, 30, 200, 640, "ARNDCQEGHILKMFPSTWYV")
    seqs += [random_seqs(rng, 1, L, L, "ARNDCQEGHILKMFPSTWYV")[0] for L in (1025, 1300, 1600, 2048, 2049)]
    check_stats(seqs)


def test_int16_range_boundary_of_the_packed_kernels():
    # slanted scores of identical tryptophan runs reach 11*L + 2*L*4 (+ go): just inside and just outside the int16 guard
    seqs = ["W" * 1600, "W" * 1680, "W" * 1690, "W" * 1450 + "A" * 200, "WY" * 700]
    check_stats(seqs)
    check_stats(seqs[:3], "BLOSUM62", 0, 4)
    # high gap extension: the slant grows faster than the scores
    seqs = ["AC" * 300, "AC" * 290 + "D" * 40, "C" * 610]
    check_stats(seqs, "BLOSUM62", 10, 25)
    check_stats(seqs, "BLOSUM62", 10, 26)


def test_nw_plan_class_runs_twice_and_slices():
    rng = np.random.default_rng(17)
    seqs = random_seqs(rng, 70, 0, 420, "ARNDCQEGHILKMFPSTWYV")
    want_m, want_l = port.nw_pair_stats(seqs, "BLOSUM80", 7, 2)
    plan = da.NWPlan(seqs, "BLOSUM80", 7, 2)
    assert plan.pairs == len(want_m) and plan.cells == sum(len(a) * len(b) for i, a in enumerate(seqs) for b in seqs[i:])
    for _ in range(2):  # the plan is reusable
        m, l = plan.run().fetch()
        assert (m == want_m).all() and (l == want_l).all()
    plan.close()
    part = da.NWPlan(seqs, "BLOSUM80", 7, 2, row_begin=20, row_end=33)
    m, l = part.run().fetch()
    wm, wl = port.nw_pair_stats(seqs, "BLOSUM80", 7, 2, row_begin=20, row_end=33)
    assert (m == wm).all() and (l == wl).all()
    with pytest.raises(da.DynaAlignError, match="Invalid substitution matrix name: PAM250"):
        da.NWPlan(seqs, "PAM250")
