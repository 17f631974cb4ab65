"""GPU parity of the result-assembly entry points (gather.cu): the tiled column-block expansion to the reference's
n x n double matrix, the position-weighted slab checksums, and the narrow host forms (2 B/pair NW, 1 B/pair MinHash
with escapes) -- each against the oracle or against the wide form it must reproduce exactly."""
import ctypes as C

import os

import numpy as np
import pytest

import dynaalign_b200 as da
from conftest import random_seqs, same_matrix
from dynaalign_b200 import _lib, synth
from oracle import port

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 63, 65, 97, 130])
def test_expansion_tiles_cover_every_shape(n):
    # tile edges of the 32 x 32 expansion (below / above / across the diagonal, ragged last tiles), NaN kept
    rng = np.random.default_rng(n)
    seqs = random_seqs(rng, n, 0, 24)
    if n > 2:
        seqs[1] = ""
        seqs[n - 1] = ""
    assert same_matrix(da.similarityNW(seqs), port.similarityNW(seqs))
    peps = random_seqs(rng, n, 0, 9, "ACDE")
    assert same_matrix(da.similarityMH(peps, 2, 17, seed=3), port.similarityMH(peps, 2, 17, 3))


def test_nw_checksum_is_additive_over_row_blocks():
    seqs = [s.decode() for s in synth.proteins_families(90)] + ["", "ACD"]
    n = len(seqs)
    wm, wl = port.nw_pair_stats(seqs)
    want = (da.checksum(wm), da.checksum(wl))
    whole = da.NWPlan(seqs)
    whole.run()
    assert whole.checksum() == want
    whole.close()
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    b = da.partition_rows(n, 3, weights=lens, include_diagonal=True)
    parts = []
    for g in range(3):
        p = da.NWPlan(seqs, row_begin=int(b[g]), row_end=int(b[g + 1]))
        p.run()
        parts.append(p.checksum())
        p.close()
    mask = (1 << 64) - 1
    assert (sum(x[0] for x in parts) & mask, sum(x[1] for x in parts) & mask) == want
    # a partition that drops one row does not add up
    p = da.NWPlan(seqs, row_begin=int(b[1]) + 1, row_end=int(b[2]))
    p.run()
    broken = [parts[0], p.checksum(), parts[2]]
    p.close()
    assert (sum(x[0] for x in broken) & mask, sum(x[1] for x in broken) & mask) != want


def test_mh_checksum_matches_oracle_counts():
    peps = [s.decode() for s in synth.peptides_clustered(700, children=10)]
    seeds = port.hashfamily_seeds(42, 60)
    want = port.mh_match_counts(port.mh_signatures(peps, 4, seeds))
    plan = da.MinHashPlan(peps, 4, 60, seeds=seeds)
    assert plan.checksum() == da.checksum(want)
    plan.close()


def test_nw_packed8_equals_wide_form():
    rng = np.random.default_rng(8)
    peps = random_seqs(rng, 400, 0, 20) + ["", "A"]
    wm, wl = port.nw_pair_stats(peps)
    m8, l8 = da.nw_pair_stats8(peps)
    assert m8.dtype == np.uint8 and (m8 == wm).all() and (l8 == wl).all()
    m8, l8 = da.nw_pair_stats8(peps, row_begin=100, row_end=317)  # a slab whose first pair is not 16-byte aligned
    wm, wl = port.nw_pair_stats(peps, row_begin=100, row_end=317)
    assert (m8 == wm).all() and (l8 == wl).all()
    with pytest.raises(da.DynaAlignError) as e:
        da.nw_pair_stats8(peps + ["A" * 200])
    assert e.value.code == _lib.ERR_UNSUPPORTED


@pytest.mark.parametrize("blocks", [2, 5, 64])
def test_nw_packed8_pipelined_row_blocks(blocks):
    # large row ranges are aligned block by block while the previous block is narrowed and copied (forced here on a small input)
    rng = np.random.default_rng(80 + blocks)
    peps = random_seqs(rng, 300, 0, 24) + ["", "", "AAAA"]
    wm, wl = port.nw_pair_stats(peps)
    os.environ["DYNA_NW_STATS8_BLOCKS"] = str(blocks)
    try:
        m8, l8 = da.nw_pair_stats8(peps)
        assert (m8 == wm).all() and (l8 == wl).all()
        m8, l8 = da.nw_pair_stats8(peps, row_begin=37, row_end=250)
        wm, wl = port.nw_pair_stats(peps, row_begin=37, row_end=250)
        assert (m8 == wm).all() and (l8 == wl).all()
        m8, l8 = da.nw_pair_stats8(peps, row_begin=301, row_end=303)   # fewer rows than blocks
        wm, wl = port.nw_pair_stats(peps, row_begin=301, row_end=303)
        assert (m8 == wm).all() and (l8 == wl).all()
        with pytest.raises(da.DynaAlignError) as e:
            da.nw_pair_stats8(peps + ["A" * 200])
        assert e.value.code == _lib.ERR_UNSUPPORTED
    finally:
        del os.environ["DYNA_NW_STATS8_BLOCKS"]


def test_mh_fetch8_with_escapes_is_lossless():
    # duplicates give count == n_hash = 300 >= 255: they must come back through the escape list
    base = [s.decode() for s in synth.peptides_clustered(900, children=30)]
    peps = base + base[:40] + ["", ""]
    seeds = port.hashfamily_seeds(7, 300)
    want = port.mh_match_counts(port.mh_signatures(peps, 4, seeds))
    plan = da.MinHashPlan(peps, 4, 300, seeds=seeds)
    c8, ei, ec = plan.match_counts8()
    plan.close()
    assert (want >= 255).sum() == ei.size and ei.size >= 41
    full = c8.astype(np.uint16)
    assert (full[ei] == 255).all()
    full[ei] = ec
    assert (full == want).all()
    # too small an escape buffer is an error that reports the number needed
    plan = da.MinHashPlan(peps, 4, 300, seeds=seeds)
    with pytest.raises(da.DynaAlignError):
        plan.match_counts8(esc_capacity=3)
    plan.close()
