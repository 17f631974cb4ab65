"""GPU parity: MinHash match counts by the join on equal signature values (csrc/mh_sparse.cu) against the oracle's
all-pairs counts (src/minHash.cpp:164-177 restated) and against the all-pairs kernel: identical histogram, threshold,
edge list, checksum and dense triangle."""
import numpy as np
import pytest

import dynaalign_b200 as da
from dynaalign_b200 import synth
from oracle import port
from oracle.quantile_r import quantile_type7

pytestmark = pytest.mark.gpu


def triangle(sig):
    return port.mh_match_counts(sig)


def make_plan(seqs, k, n_hash, join, seed=42):
    plan = da.MinHashPlan(seqs, k, n_hash, seed=seed)
    plan.join = join
    return plan


@pytest.fixture()
def pack16(monkeypatch):
    # small inputs only get sorted hash rows (and with them the join) when the 16-bit relabelling is forced on
    monkeypatch.setenv("DYNA_MH_PACK16", "1")


def mixed_peptides(n, seed):
    rng = np.random.default_rng(seed)
    seqs = [s.decode() for s in synth.peptides_clustered(n - 40, children=7)]
    seqs += ["ACD", "AC", ""] * 5              # shorter than k: all-UINT32_MAX signatures, one big group in every hash row
    seqs += [seqs[3]] * 13 + [seqs[17]] * 12   # exact duplicates: groups in every hash row
    order = rng.permutation(len(seqs))
    return [seqs[t] for t in order]


@pytest.mark.parametrize("n,k,n_hash", [(300, 4, 50), (1100, 3, 31), (2500, 4, 120)])
def test_join_equals_all_pairs(pack16, n, k, n_hash):
    seqs = mixed_peptides(n, n)
    seeds = port.hashfamily_seeds(42, n_hash)
    want = triangle(port.mh_signatures(seqs, k, seeds))
    dense = make_plan(seqs, k, n_hash, "never")
    sparse = make_plan(seqs, k, n_hash, "always")
    assert sparse.match_sparse(0) and sparse.joined
    assert sparse.incidences == int(want.astype(np.int64).sum())  # one incidence per matching (pair, hash function)
    hist = np.bincount(want, minlength=n_hash + 1).astype(np.uint64)
    assert (sparse.histogram() == hist).all() and (dense.histogram() == hist).all()
    assert sparse.checksum() == dense.checksum() == da.checksum(want)
    for p in (0.0, 0.5, 0.8, 0.97, 1.0):
        a, b = sparse.threshold_edges(p), dense.threshold_edges(p)
        assert a[0] == b[0] == quantile_type7(want / n_hash, p)
        for x, y in zip(a[1:], b[1:]):
            assert len(x) == len(y) and (x == y).all()
    assert (sparse.match_counts() == want).all()  # scattered into the dense triangle on the device
    dense.close()
    sparse.close()


def test_join_respects_row_ranges(pack16):
    import ctypes as C
    from dynaalign_b200 import _lib
    seqs = mixed_peptides(900, 5)
    n, n_hash, k = len(seqs), 40, 4
    seeds = port.hashfamily_seeds(7, n_hash)
    want = triangle(port.mh_signatures(seqs, k, seeds))
    L = _lib.lib()
    res, off = _lib.flatten(seqs)
    b = da.partition_rows(n, 3)
    hist, got = np.zeros(n_hash + 1, dtype=np.uint64), []
    for s in range(3):
        plan = L.dyna_mh_plan_create(n, n_hash, int(b[s]), int(b[s + 1]), 0)
        assert plan, _lib.last_error()
        try:
            _lib.check(L.dyna_mh_plan_upload_sequences(plan, _lib.ptr(res, C.c_uint8), _lib.ptr(off, C.c_int64), k, _lib.ptr(seeds, C.c_uint32), None))
            _lib.check(L.dyna_mh_plan_run_signatures(plan, None))
            inc, done = C.c_int64(0), C.c_int(0)
            _lib.check(L.dyna_mh_plan_run_match_sparse(plan, 0, C.byref(inc), C.byref(done), None))
            assert done.value == 1
            h = np.zeros(n_hash + 1, dtype=np.uint64)
            _lib.check(L.dyna_mh_plan_count_histogram(plan, _lib.ptr(h, C.c_uint64), None))
            hist += h
            part = np.zeros(max(L.dyna_mh_plan_pairs(plan), 1), dtype=np.uint16)
            _lib.check(L.dyna_mh_plan_fetch_counts(plan, _lib.ptr(part, C.c_uint16), None))
            got.append(part[:L.dyna_mh_plan_pairs(plan)])
        finally:
            L.dyna_mh_plan_destroy(plan)
    assert (np.concatenate(got) == want).all()
    assert (hist == np.bincount(want, minlength=n_hash + 1).astype(np.uint64)).all()


def test_join_declines_when_capped_or_unsorted(pack16, monkeypatch):
    seqs = mixed_peptides(400, 9)
    plan = make_plan(seqs, 4, 30, "auto")
    assert not plan.match_sparse(10)  # far more matches than 10: the caller falls back
    assert plan.incidences > 10
    thr, ei, ej, w = plan.threshold_edges(0.8)  # auto: cap from the cost model, then all-pairs if it declines
    ref = make_plan(seqs, 4, 30, "never").threshold_edges(0.8)
    assert thr == ref[0] and (ei == ref[1]).all() and (ej == ref[2]).all() and (w == ref[3]).all()
    plan.close()
    monkeypatch.setenv("DYNA_MH_PACK16", "0")  # no relabelling, no sorted rows: the join is not available
    plan = make_plan(seqs, 4, 30, "always")
    assert not plan.match_sparse(0) and plan.incidences is None
    assert (plan.histogram() == make_plan(seqs, 4, 30, "never").histogram()).all()
    plan.close()


def test_join_without_a_dense_triangle_for_large_n():
    # 60,000 random 16-mers: 1.8e9 pairs.  The join never allocates the 3.6 GB triangle; results equal the all-pairs kernel's
    seqs = [s.decode() for s in synth.peptides_uniform(60000)]
    sparse = make_plan(seqs, 4, 64, "auto")
    thr, ei, ej, w = sparse.threshold_edges(0.999)
    assert sparse.joined and sparse.incidences < 10 ** 7
    dense = make_plan(seqs, 4, 64, "never")
    ref = dense.threshold_edges(0.999)
    assert not dense.joined
    assert thr == ref[0] and len(ei) == len(ref[1]) and (ei == ref[1]).all() and (ej == ref[2]).all() and (w == ref[3]).all()
    assert sparse.checksum() == dense.checksum()
    assert (sparse.histogram() == dense.histogram()).all()
    dense.close()
    sparse.close()


def test_subset_plans_join_too(pack16):
    seqs = mixed_peptides(700, 21)
    plan = make_plan(seqs, 4, 50, "always")
    rng = np.random.default_rng(3)
    idx = np.sort(rng.choice(len(seqs), size=333, replace=False))
    sub = plan.subset(idx)
    seeds = port.hashfamily_seeds(42, 50)
    want = triangle(port.mh_signatures([seqs[t] for t in idx], 4, seeds))
    assert (sub.histogram() == np.bincount(want, minlength=51).astype(np.uint64)).all() and sub.joined
    assert (sub.match_counts() == want).all()
    sub.close()
    plan.close()
