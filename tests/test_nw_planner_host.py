"""Host-side NW planner (no GPU): the work units of all kernel classes together cover every pair (i <= j) of the row
range exactly once -- for homogeneous inputs, proteome-like length mixes, long outliers, empty sequences, row slabs and
penalty settings that leave the 16-bit range.  (dyna_nw_plan_layout runs the same code as dyna_nw_plan_create up to the
point where device memory is allocated.)"""
import ctypes as C

import numpy as np
import pytest

from dynaalign_b200._lib import flatten, lib, ptr

AL = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)
TWO_ROW_KINDS = (8, 9, 10, 12)


def layout(lens, table=b"BLOSUM62", go=10, ge=4, rb=0, re_=None, seed=0):
    rng = np.random.default_rng(seed)
    seqs = [AL[rng.integers(0, 20, size=int(L))].tobytes() for L in lens]
    res, off = flatten(seqs)
    n = len(seqs)
    re_ = n if re_ is None else re_
    L = lib()
    p = L.dyna_nw_plan_layout(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, table, go, ge, rb, re_)
    assert p
    try:
        k = L.dyna_nw_plan_unit_count(p)
        out = np.zeros((max(k, 1), 6), dtype=np.int32)
        assert L.dyna_nw_plan_export_units(p, ptr(out, C.c_int32)) == 0
        pairs = L.dyna_nw_plan_pairs(p)
    finally:
        L.dyna_nw_plan_destroy(p)
    return out[:k], pairs


def check_cover(lens, rb=0, re_=None, **kw):
    lens = np.asarray(lens)
    n = len(lens)
    re_ = n if re_ is None else re_
    units, pairs = layout(lens, rb=rb, re_=re_, **kw)
    cover = np.zeros((n, n), dtype=np.int32)
    for kind, R, row, row2, j0, cnt in units:
        assert rb <= row < re_ and j0 >= row and j0 + cnt <= n and cnt >= 1
        cover[row, j0:j0 + cnt] += 1
        if kind in TWO_ROW_KINDS:
            assert row < row2 < re_
            lo = max(j0, row2)
            if lo < j0 + cnt:
                cover[row2, lo:j0 + cnt] += 1
        else:
            assert row2 == -1
    want = np.zeros((n, n), dtype=np.int32)
    for i in range(rb, re_):
        want[i, i:] = 1
    assert (cover == want).all(), "pairs covered %d times: %s" % (cover[cover != want][0], np.argwhere(cover != want)[:5])
    assert pairs == int(want.sum())
    return units


def test_homogeneous_proteins_pair_by_length_inside_the_window():
    rng = np.random.default_rng(1)
    lens = np.clip(np.rint(rng.normal(330, 10, size=1200)), 300, 360).astype(int)
    units = check_cover(lens)
    assert set(units[:, 0]) == {8}, "every row of a homogeneous input goes through the two-rows kernel"
    d = units[:, 3] - units[:, 2]
    assert (d >= 1).all() and (d <= 32).all()
    # partners are chosen by length inside the window (n / 96 = 12 rows here)
    assert np.abs(lens[units[:, 3]] - lens[units[:, 2]]).mean() < 6  # neighbours would differ by ~11


def test_mixed_lengths_pair_inside_the_window():
    rng = np.random.default_rng(2)
    lens = np.clip(rng.lognormal(np.log(300.0), 0.5, size=1000), 30, 1800).astype(int)
    units = check_cover(lens)
    two = units[np.isin(units[:, 0], TWO_ROW_KINDS)]
    assert len(two) > 0 and (two[:, 3] - two[:, 2]).max() > 1, "partners beyond the neighbour"
    assert (two[:, 3] - two[:, 2]).max() <= 32


def test_long_outlier_only_takes_its_own_units_off_the_two_rows_kernel():
    lens = [330] * 600
    lens[450] = 2500                      # beyond the staging buffer of the two-rows kernel
    lens[100] = 1500                      # inside it
    units = check_cover(lens)
    k8 = units[units[:, 0] == 8]
    # rows before the outlier keep the two-rows kernel for the column blocks that do not contain it
    assert ((k8[:, 2] < 440) & (k8[:, 4] + k8[:, 5] <= 450)).any() and ((k8[:, 2] < 440) & (k8[:, 4] > 450)).any()
    assert not ((k8[:, 4] <= 450) & (k8[:, 4] + k8[:, 5] > 450)).any(), "no two-rows unit contains the 2500-residue column"
    assert ((k8[:, 4] <= 100) & (k8[:, 4] + k8[:, 5] > 100)).any(), "the 1500-residue column stays in two-rows units"


@pytest.mark.parametrize("seed", range(6))
def test_random_mixes_and_row_slabs(seed):
    rng = np.random.default_rng(100 + seed)
    kind = seed % 3
    if kind == 0:
        lens = rng.integers(0, 900, size=150)
    elif kind == 1:
        lens = np.concatenate([rng.integers(1, 33, size=120), rng.integers(300, 1500, size=6), [0, 0]])
    else:
        lens = np.concatenate([rng.integers(385, 769, size=90), rng.integers(30, 400, size=40), [2049, 3000]])
    rng.shuffle(lens)
    check_cover(lens, seed=seed)
    n = len(lens)
    a, b = sorted(rng.integers(0, n + 1, size=2).tolist())
    check_cover(lens, rb=a, re_=b, seed=seed)


def test_penalties_outside_the_16_bit_range_use_the_32_bit_kernels():
    lens = [200, 210, 220, 500, 510, 20, 22]
    units = check_cover(lens, go=9000, ge=4)
    assert not set(units[:, 0]) & {4, 5, 6, 7, 8, 9, 10, 12}
    check_cover(lens, go=0, ge=0)
    check_cover(lens, table=b"BLOSUM45", go=3, ge=1)


def test_long_rows_take_the_two_rows_multipass_kernel():
    rng = np.random.default_rng(9)
    lens = np.concatenate([rng.integers(800, 1000, size=40), rng.integers(1200, 1500, size=20), rng.integers(100, 300, size=30)])
    rng.shuffle(lens)
    units = check_cover(lens)
    k12 = units[units[:, 0] == 12]
    assert len(k12) > 0
    R = k12[:, 1]
    for kind, r, row, row2, j0, cnt in k12:
        a, b = sorted((int(lens[row]), int(lens[row2])))
        assert 7 <= r <= 12 and (a - 1) // (32 * r) == (b - 1) // (32 * r), "both rows end in the same pass"


def test_few_long_rows_go_back_to_the_one_row_multipass_kernel():
    # the two-rows multi-pass kernel is used only when its rows hold a tenth of the plan's cells
    rng = np.random.default_rng(10)
    lens = np.concatenate([rng.integers(300, 360, size=600), rng.integers(800, 1000, size=6)])
    rng.shuffle(lens)
    units = check_cover(lens)
    assert 12 not in set(units[:, 0]) and 6 in set(units[:, 0])


def test_small_plans_get_narrow_two_rows_units():
    lens = np.full(300, 330)
    units = check_cover(lens)
    k8 = units[units[:, 0] == 8]
    assert len(k8) > 0 and k8[:, 5].max() <= 32      # 45,150 pairs: 32-column units
    units = check_cover(np.full(1000, 330))
    assert units[units[:, 0] == 8][:, 5].max() == 128   # 500,500 pairs: 128-column units
