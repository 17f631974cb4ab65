"""GPU, BASELINE.json's full sizes: the oracle cannot run 2e13 DP cells or 5e9 pairs, so the full-size runs are
checked through size-independent properties plus a random sample of pairs recomputed with the oracle."""
import ctypes as C

import numpy as np
import pytest

import dynaalign_b200 as da
from dynaalign_b200 import _lib, synth
from dynaalign_b200._lib import check, flatten, lib, ptr
from oracle import port

pytestmark = pytest.mark.gpu


def test_config5_full_nw_20k_proteins():
    # synthetic 20k proteins of ~330 aa, all 200,010,000 pairs (2.2e13 cells)
    seqs = synth.proteins_families(20000)
    n = len(seqs)
    L = lib()
    res, off = flatten(seqs)
    lens = np.diff(off)
    plan = L.dyna_nw_plan_create(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, b"BLOSUM62", 10, 4, 0, n, 0)
    assert plan, _lib.last_error()
    try:
        pairs = L.dyna_nw_plan_pairs(plan)
        assert pairs == n * (n + 1) // 2
        suffix = np.cumsum(lens[::-1])[::-1]
        assert L.dyna_nw_plan_cells(plan) == int((lens * suffix).sum())
        check(L.dyna_nw_plan_run(plan, None))
        mt = np.zeros(pairs, dtype=np.uint32)
        ln = np.zeros(pairs, dtype=np.uint32)
        check(L.dyna_nw_plan_fetch(plan, ptr(mt, C.c_uint32), ptr(ln, C.c_uint32), None))
    finally:
        L.dyna_nw_plan_destroy(plan)
    idx = lambda i, j: i * n - i * (i - 1) // 2 + (j - i)
    # 1. self-alignments are identities: matches == length == len(seq)
    diag = np.array([idx(i, i) for i in range(n)])
    assert (mt[diag] == lens).all() and (ln[diag] == lens).all()
    # 2. global bounds: max(m,n) <= length <= m+n, matches <= min(m,n), matches <= length
    rng = np.random.default_rng(1)
    ii = rng.integers(0, n, 200000)
    jj = rng.integers(0, n, 200000)
    i_, j_ = np.minimum(ii, jj), np.maximum(ii, jj)
    k = i_ * n - i_ * (i_ - 1) // 2 + (j_ - i_)
    assert (ln[k] >= np.maximum(lens[i_], lens[j_])).all() and (ln[k] <= lens[i_] + lens[j_]).all()
    assert (mt[k] <= np.minimum(lens[i_], lens[j_])).all()
    # 3. a random sample of pairs (within-family and across families) recomputed by the oracle
    sample = [(int(a), int(b)) for a, b in zip(i_[:120], j_[:120])]
    sample += [(f * 100 + int(a), f * 100 + int(b)) for f in (0, 57, 199) for a, b in [sorted(rng.integers(0, 100, 2)) for _ in range(20)]]
    for a, b in sample:
        want = port.nw_pair(seqs[a], seqs[b])
        assert (int(mt[idx(a, b)]), int(ln[idx(a, b)])) == want, (a, b)
    # 4. checksum of the whole result is stable across row-block sharding (first and last blocks re-run separately)
    b = da.partition_rows(n, 8, weights=lens, include_diagonal=True)
    for s in (0, 7):
        m2, l2 = da.nw_pair_stats(seqs, row_begin=int(b[s]), row_end=int(b[s + 1]))
        lo, hi = idx(int(b[s]), int(b[s])), (idx(int(b[s + 1]), int(b[s + 1])) if b[s + 1] < n else pairs)
        assert (m2 == mt[lo:hi]).all() and (l2 == ln[lo:hi]).all()


def test_config4_full_minhash_100k_peptides():
    # synthetic 100k peptides of 16 aa, k=4, n_hash=500: 4,999,950,000 pairs; checked on device-side reductions
    n, n_hash, k = 100000, 500, 4
    seqs = synth.peptides_clustered(n)  # the clustered variant has real matches
    L = lib()
    res, off = flatten(seqs)
    seeds = da.hashfamily_seeds(42, n_hash)
    plan = L.dyna_mh_plan_create(n, n_hash, 0, n, 0)
    assert plan, _lib.last_error()
    try:
        check(L.dyna_mh_plan_upload_sequences(plan, ptr(res, C.c_uint8), ptr(off, C.c_int64), k, ptr(seeds, C.c_uint32), None))
        check(L.dyna_mh_plan_run_signatures(plan, None))
        check(L.dyna_mh_plan_run_match(plan, None))
        sig = np.zeros((n, n_hash), dtype=np.uint32)
        check(L.dyna_mh_plan_fetch_signatures(plan, ptr(sig, C.c_uint32), None))
        hist = np.zeros(n_hash + 1, dtype=np.uint64)
        check(L.dyna_mh_plan_count_histogram(plan, ptr(hist, C.c_uint64), None))
        # edges with >= 100 matches (the family structure): small enough to fetch and verify one by one
        cap = int(hist[100:].sum())
        ei = np.zeros(cap, dtype=np.int32)
        ej = np.zeros(cap, dtype=np.int32)
        ec = np.zeros(cap, dtype=np.uint16)
        ne = C.c_int64(0)
        check(L.dyna_mh_plan_threshold_edges(plan, 100, cap, ptr(ei, C.c_int32), ptr(ej, C.c_int32), ptr(ec, C.c_uint16), C.byref(ne), None))
    finally:
        L.dyna_mh_plan_destroy(plan)
    # 1. signatures of a sample of sequences against the oracle (idempotent re-hash)
    pick = np.random.default_rng(2).integers(0, n, 300)
    assert (sig[pick] == port.mh_signatures([seqs[i] for i in pick], k, seeds)).all()
    # 2. every pair is counted exactly once
    assert int(hist.sum()) == n * (n - 1) // 2
    # 3. checksum of checksums: sum over pairs of matches == sum over hash rows of sum_v C(multiplicity(v), 2)
    total = 0
    for h in range(n_hash):
        _, mult = np.unique(sig[:, h], return_counts=True)
        total += int((mult.astype(np.int64) * (mult - 1) // 2).sum())
    assert int((hist * np.arange(n_hash + 1, dtype=np.uint64)).sum()) == total
    # 4. the high-count edges, one by one against the signatures
    assert ne.value == cap and cap > 10000
    sel = np.random.default_rng(3).integers(0, cap, 20000)
    got = (sig[ei[sel]] == sig[ej[sel]]).sum(axis=1)
    assert (got == ec[sel]).all() and (ei[sel] < ej[sel]).all()
    assert (np.diff(ei.astype(np.int64) * n + ej) > 0).all()  # row-major, strictly increasing


def test_target_full_nw_100k_peptides():
    # BASELINE.json's stated target: the N x N NW result for the 100,000 config-4 peptides -- 5,000,050,000 pairs
    # (more than 2^32: every pair index has to be 64-bit), 1.28e12 cells, one GPU
    seqs = synth.peptides_uniform(100000)
    n = len(seqs)
    L = lib()
    res, off = flatten(seqs)
    plan = L.dyna_nw_plan_create(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, b"BLOSUM62", 10, 4, 0, n, 0)
    assert plan, _lib.last_error()
    try:
        pairs = L.dyna_nw_plan_pairs(plan)
        assert pairs == n * (n + 1) // 2 and L.dyna_nw_plan_cells(plan) == pairs * 256
        check(L.dyna_nw_plan_run(plan, None))
        mt = np.zeros(pairs, dtype=np.uint32)
        ln = np.zeros(pairs, dtype=np.uint32)
        check(L.dyna_nw_plan_fetch(plan, ptr(mt, C.c_uint32), ptr(ln, C.c_uint32), None))
    finally:
        L.dyna_nw_plan_destroy(plan)
    idx = lambda i, j: i * n - i * (i - 1) // 2 + (j - i)
    rows = np.arange(n, dtype=np.int64)
    diag = rows * n - rows * (rows - 1) // 2
    assert (mt[diag] == 16).all() and (ln[diag] == 16).all()
    assert int(ln.min()) >= 16 and int(ln.max()) <= 32 and int(mt.max()) <= 16 and (mt <= ln).all()
    rng = np.random.default_rng(5)
    ii, jj = rng.integers(0, n, 1500), rng.integers(0, n, 1500)
    sample = list(zip(np.minimum(ii, jj).tolist(), np.maximum(ii, jj).tolist()))
    sample += [(0, n - 1), (n - 2, n - 1), (n - 1, n - 1), (61000, 61001), (99999 - 7, 99999)]
    for a, b in sample:
        assert (int(mt[idx(a, b)]), int(ln[idx(a, b)])) == port.nw_pair(seqs[a], seqs[b]), (a, b)
    # a row block that straddles pair index 2^32, re-run through the host-buffer entry point
    r0 = int(np.searchsorted(diag, 1 << 32)) - 3
    m2, l2 = da.nw_pair_stats(seqs, row_begin=r0, row_end=r0 + 6)
    lo = idx(r0, r0)
    assert lo < (1 << 32) < lo + len(m2)
    assert (m2 == mt[lo:lo + len(m2)]).all() and (l2 == ln[lo:lo + len(l2)]).all()
