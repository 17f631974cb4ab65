"""GPU parity: MinHash kernels (through the C ABI) against the oracle, bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

import dynaalign_b200 as da
from conftest import GOLDEN, fingerprint, random_seqs, same_matrix
from dynaalign_b200 import _lib
from oracle import minhash_r as R
from oracle import port

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,n_hash", [(1, 7), (2, 50), (3, 33), (4, 500), (5, 64), (7, 129), (8, 16), (9, 20), (13, 5)])
def test_signatures_bit_exact(k, n_hash):
    rng = np.random.default_rng(100 + k)
    seqs = random_seqs(rng, 60, 0, 40, "ACDEFGHIKLMNPQRSTVWY") + ["", "A", "AC", "ACDEFGHIKLMN" * 30]
    seeds = port.hashfamily_seeds(9 + k, n_hash)
    got = da.mh_signatures(seqs, k, seeds)
    want = port.mh_signatures(seqs, k, seeds)
    assert (got == want).all()
    # sequences shorter than k keep UINT32_MAX (src/minHash.cpp:140)
    assert (got[60] == 0xFFFFFFFF).all()


def test_signatures_warp_reduction_path():
    # few hash functions on long sequences take the lanes-over-windows + redux.min kernel
    rng = np.random.default_rng(3)
    seqs = random_seqs(rng, 20, 300, 2500, "ACDEFGHIKLMNPQRSTVWY")
    seeds = port.hashfamily_seeds(77, 12)
    assert (da.mh_signatures(seqs, 5, seeds) == port.mh_signatures(seqs, 5, seeds)).all()


def test_signatures_raw_bytes_and_long_k():
    seqs = [bytes(range(1, 200)), b"\xff\xfe\x00abc" * 9, b"lower case too", b"x"]
    seeds = port.hashfamily_seeds(5, 40)
    for k in (3, 4, 6, 31, 64):
        assert (da.mh_signatures(seqs, k, seeds) == port.mh_signatures(seqs, k, seeds)).all()


@pytest.mark.parametrize("n,n_hash", [(2, 1), (3, 50), (129, 17), (257, 500), (700, 64), (1000, 33)])
@pytest.mark.parametrize("mode", ["tma", "ldg"])
def test_match_counts_bit_exact(n, n_hash, mode, monkeypatch):
    monkeypatch.setenv("DYNA_MH_MATCH", mode)
    rng = np.random.default_rng(n * 31 + n_hash)
    sig = rng.integers(0, 3, size=(n, n_hash), dtype=np.uint32) * np.uint32(0x9E3779B1)  # many collisions, wide values
    sig[rng.integers(0, n, 3)] = 0xFFFFFFFF
    assert (da.mh_match_counts(sig) == port.mh_match_counts(sig)).all()


def test_match_counts_row_slabs():
    rng = np.random.default_rng(8)
    n = 900
    sig = rng.integers(0, 5, size=(n, 100), dtype=np.uint32)
    full = port.mh_match_counts(sig)
    bounds = da.partition_rows(n, 5)
    parts = [da.mh_match_counts(sig, int(bounds[s]), int(bounds[s + 1])) for s in range(5)]
    assert (np.concatenate(parts) == full).all()
    # unaligned row ranges
    for a, b in [(1, 2), (127, 130), (333, 899), (899, 900), (5, 5)]:
        assert (da.mh_match_counts(sig, a, b) == port.mh_match_counts(sig, a, b)).all()


def test_similarityMH_evp_config1(golden, evp):
    # BASELINE config 1: similarityMH(evp_peparray$PROBE_SEQUENCE, k=2, n_hash=50), seed injected
    m = da.similarityMH(evp, 2, 50, seed=42)
    assert fingerprint(m) == golden["mh_evp_k2_h50_seed42"]["fnv1a64"]
    assert same_matrix(m, port.similarityMH(evp, 2, 50, 42))
    sig = np.load(os.path.join(GOLDEN, "mh_evp_signatures_k2_h50_seed42.npz"))["sig"]
    assert (da.mh_signatures(evp, 2, da.hashfamily_seeds(42, 50)) == sig).all()
    assert (np.diag(m) == 1.0).all() and (m == m.T).all()


def test_similarityMH_h3n2_config3_input(golden, h3n2):
    m = da.similarityMH(h3n2, 4, 500, seed=42)
    assert fingerprint(m) == golden["mh_h3n2_1000_k4_h500_seed42"]["fnv1a64"]


def test_similarityMH_edge_cases():
    # one sequence; all shorter than k (mutually identical: sim 1.0); duplicates
    assert da.similarityMH(["ACDEF"], 4, 10, seed=1).tolist() == [[1.0]]
    m = da.similarityMH(["AC", "A", "", "ACDEFG", "ACDEFG"], 4, 20, seed=3)
    assert same_matrix(m, port.similarityMH(["AC", "A", "", "ACDEFG", "ACDEFG"], 4, 20, 3))
    assert m[0, 1] == 1.0 and m[0, 2] == 1.0 and m[3, 4] == 1.0 and m[0, 3] == 0.0


def test_similarityMH_default_seed_is_random_but_valid(evp):
    a = da.similarityMH(evp[:50], 2, 50)
    assert (np.diag(a) == 1.0).all() and (a == a.T).all() and ((a >= 0) & (a <= 1)).all()
    assert np.allclose(a * 50, np.round(a * 50))


def test_r_pipeline_on_gpu():
    rng = np.random.default_rng(4)
    seqs = random_seqs(rng, 40, 6, 30, "ACDEGHIKLMN")
    vocab = R.create_vocab(seqs, 3)
    assert da.create_vocab(seqs, 3) == vocab
    cm = da.create_char_matrix(seqs, vocab, 3)
    hp = R.create_hash_parameters(64, len(vocab), np.random.default_rng(5))
    want_sig = R.compute_signature_matrix(cm, hp, len(vocab))
    got_sig = da.compute_signature_matrix(cm, hp, len(vocab))
    assert got_sig.shape == (64, 40) and (got_sig == want_sig).all()
    want_d = R.compute_distance_matrix(want_sig)
    got_d = da.compute_distance_matrix(got_sig)
    assert same_matrix(got_d, want_d)
    r = da.minhash(seqs, 3, 64, hash_params=hp)
    assert same_matrix(r["dist_matrix"], want_d) and r["vocabulary"] == vocab
    # testthat mock (tests/testthat/test-minHash.R:92-106)
    mock = np.array([[1, 2, 3], [1, 2, 4], [2, 3, 5]], dtype=float).T
    d = da.compute_distance_matrix(mock)
    assert d[0, 1] == 1 - 2 / 3 and d[0, 2] == 1 and d[1, 2] == 1 and (np.diag(d) == 0).all()


def test_linear_signatures_large_values():
    # stay inside R's integer domain (a*x+b < 2^31) and also beyond it (64-bit on both sides)
    rng = np.random.default_rng(6)
    for m in (46340, 2_000_000_011):
        n = 25
        counts = rng.integers(0, 40, n)
        off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        ranks = rng.integers(1, min(m, 2 ** 31 - 1), size=int(off[-1]) or 1).astype(np.int32)
        a = rng.integers(1, m, 30).astype(np.int64)
        b = rng.integers(0, m, 30).astype(np.int64)
        want = port.mh_signatures_linear(ranks, off, a, b, m)
        got = np.zeros((n, 30), dtype=np.uint32)
        _lib.check(_lib.lib().dyna_mh_signatures_linear(_lib.ptr(ranks, C.c_int32), _lib.ptr(off, C.c_int64), n,
                                                        _lib.ptr(a, C.c_int64), _lib.ptr(b, C.c_int64), m, 30,
                                                        _lib.ptr(got, C.c_uint32)))
        assert (got == want).all()


@pytest.mark.parametrize("n,n_hash", [(2, 1), (130, 7), (300, 50), (700, 501), (2500, 100)])
def test_match_counts_16bit_relabelled_path(n, n_hash, monkeypatch):
    # HSET2 path: rows relabelled to fp16-safe dense codes, two hash components per word (odd n_hash -> padded half)
    monkeypatch.setenv("DYNA_MH_PACK16", "1")
    rng = np.random.default_rng(n + n_hash)
    sig = rng.integers(0, 6, size=(n, n_hash), dtype=np.uint32) * np.uint32(0x9E3779B1)
    sig[rng.integers(0, n, 3)] = 0xFFFFFFFF
    sig[:, 0] = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)  # an (almost) all-distinct row
    want = port.mh_match_counts(sig)
    assert (da.mh_match_counts(sig) == want).all()
    monkeypatch.setenv("DYNA_MH_PACK16", "0")
    assert (da.mh_match_counts(sig) == want).all()


def test_16bit_path_overflow_falls_back_exactly(monkeypatch):
    # more distinct values in a hash row than codes -> the gate sends the work to the 32-bit kernel
    monkeypatch.setenv("DYNA_MH_PACK16", "1")
    monkeypatch.setenv("DYNA_MH_MAXCODES", "5")
    rng = np.random.default_rng(21)
    sig = rng.integers(0, 9, size=(400, 33), dtype=np.uint32)
    assert (da.mh_match_counts(sig) == port.mh_match_counts(sig)).all()
    sig = rng.integers(0, 4, size=(400, 33), dtype=np.uint32)  # fits 5 codes: stays on the 16-bit kernel
    assert (da.mh_match_counts(sig) == port.mh_match_counts(sig)).all()


def test_similarityMH_large_default_path():
    # n >= 2048 switches the 16-bit path on by default; row slabs through it as well
    from dynaalign_b200 import synth
    seqs = [s.decode() for s in synth.peptides_clustered(3000, children=20)]
    seeds = port.hashfamily_seeds(42, 120)
    sig = da.mh_signatures(seqs, 4, seeds)
    assert (sig == port.mh_signatures(seqs, 4, seeds)).all()
    want = port.mh_match_counts(sig)
    assert (da.mh_match_counts(sig) == want).all()
    assert want.max() > 60  # the clustered set really has matching pairs
    b = da.partition_rows(3000, 3)
    parts = [da.mh_match_counts(sig, int(b[s]), int(b[s + 1])) for s in range(3)]
    assert (np.concatenate(parts) == want).all()


@pytest.mark.parametrize("p", [0.0, 0.25, 0.8, 0.9371, 1.0])
def test_threshold_and_edge_list(p, evp):
    # clusterbreak's step after sim_fn (R/clusterbreak.R:219-221) from the device-resident counts
    from oracle.quantile_r import quantile_type7
    full = port.similarityMH(evp, 2, 50, 42)
    iu = np.triu_indices(len(evp), 1)
    want_thr = quantile_type7(full[iu], p)
    thr, ei, ej, w = da.similarityMH_edges(evp, 2, 50, p, seed=42)
    assert thr == want_thr
    assert abs(thr - np.quantile(full[iu], p)) < 1e-12  # numpy's "linear" method is type 7 up to rounding
    dense = full.copy()
    dense[dense < want_thr] = 0.0  # pep.sim[pep.sim < threshold] <- 0
    wi, wj = np.nonzero(np.triu(dense, 1))
    assert (ei == wi).all() and (ej == wj).all() and (w == dense[wi, wj]).all()


def test_plan_subsets_reuse_signatures(evp):
    # clusterbreak's recursion: sim_fn on a sub-cluster == the sub-matrix of match counts, without re-hashing
    from oracle.quantile_r import quantile_type7
    plan = da.MinHashPlan(evp, 2, 50, seed=42)
    full = port.similarityMH(evp, 2, 50, 42)
    rng = np.random.default_rng(4)
    for size in (2, 37, 300):
        idx = np.sort(rng.choice(len(evp), size=size, replace=False))
        sub = plan.subset(idx)
        counts = sub.match_counts()
        want = np.rint(full[np.ix_(idx, idx)][np.triu_indices(size, 1)] * 50).astype(np.uint16)
        assert (counts == want).all()
        thr, ei, ej, w = sub.threshold_edges(0.8)
        assert thr == quantile_type7(full[np.ix_(idx, idx)][np.triu_indices(size, 1)], 0.8)
        sub.close()
    assert (plan.signatures() == port.mh_signatures(evp, 2, port.hashfamily_seeds(42, 50))).all()
    plan.close()


def test_device_vocabulary_and_full_gpu_minhash(evp):
    # create_vocab / ranks on the device == the R restatement; full pipeline == R restatement given the same (a, b)
    for seqs, k in [(evp[:200], 3), (["ACDEGHHIKLLL", "ACDEGHHIKLMN", "XXXXXYYYYYYZZ"], 3), (evp, 1), (evp[:50], 8)]:
        vocab, ranks, roff = da.vocab_ranks(seqs, k)
        assert vocab == R.create_vocab(seqs, k)
        wr, wo = R.shingle_ranks(seqs, vocab, k)
        assert (ranks == wr).all() and (roff == wo).all()
    seqs = evp[:120]
    vocab = R.create_vocab(seqs, 4)
    hp = R.create_hash_parameters(200, len(vocab), np.random.default_rng(8))
    want = R.minhash(seqs, 4, 200, hash_params=hp)
    got = da.minhash_gpu(seqs, 4, 200, hash_params=hp)
    assert got["vocabulary"] == want["vocabulary"]
    assert (got["sig_matrix"] == want["sig_matrix"]).all()
    assert same_matrix(got["dist_matrix"], want["dist_matrix"])
    with pytest.raises(da.DynaAlignError, match="'k' must be a positive integer between 1 and 2"):
        da.vocab_ranks(["ACDE", "AC", "ACDEF"], 3)


def test_overlapped_match_fetch_equals_plain():
    from dynaalign_b200 import synth
    seqs = [s.decode() for s in synth.peptides_clustered(20000, children=50)]
    plan = da.MinHashPlan(seqs, 4, 64, seed=7)
    a = plan.match_counts()          # chunked + overlapped (170 MB of counts -> several chunks)
    b = plan.match_counts()          # plain fetch of the same device buffer
    assert (a == b).all()
    sig = plan.signatures()
    pick = np.random.default_rng(0).integers(0, len(a), 5000)
    n = len(seqs)
    # invert the packed index for a sample and recount from the signatures
    i = (n - 2 - np.floor(np.sqrt(-8.0 * pick + 4.0 * n * (n - 1) - 7) / 2.0 - 0.5)).astype(np.int64)
    j = (pick + i + 1 - n * (n - 1) // 2 + (n - i) * ((n - i) - 1) // 2).astype(np.int64)
    assert ((sig[i] == sig[j]).sum(axis=1) == a[pick]).all()
    plan.close()


def _dense_clusterbreak(pep, sim_fn, cluster_fn, thresh_p, size_max, size_min, max_itr):
    """Literal dense-matrix restatement of cluster_recursive (R/clusterbreak.R:203-259) for the test below."""
    from oracle.quantile_r import quantile_type7
    state = {"rows": [], "itr": 1, "conv": 1, "filtered": []}

    def rec(seqs):
        if state["itr"] > max_itr:
            state["conv"] = 0
            return
        sim = np.array(sim_fn(seqs), dtype=np.float64)
        n = len(seqs)
        if n > 1:
            thr = quantile_type7(sim[np.triu_indices(n, 1)], thresh_p)
            sim[sim < thr] = 0.0
        gi, gj = np.nonzero(np.triu(sim))  # mode = "upper": diagonal included
        c = np.asarray(cluster_fn(n, gi, gj, sim[gi, gj]))
        size = np.bincount(c)[1:]
        ids = np.arange(1, len(size) + 1)
        big, small = ids[size > size_max], ids[size < size_min]
        state["filtered"] += [seqs[t] for t in range(n) if c[t] in small]
        label = state["itr"]
        state["rows"] += [(seqs[t], "%d.%d" % (label, c[t])) for t in range(n) if c[t] not in small and c[t] not in big]
        order = []
        for t in range(n):
            if c[t] in big and c[t] not in order:
                order.append(c[t])
        for cid in order:
            state["itr"] += 1
            rec([seqs[t] for t in range(n) if c[t] == cid])

    rec(list(pep))
    return state


def _weighted_degree_mod3(n, i, j, w):
    # test-only cluster_fn: sensitive to every edge and weight, and splits any set into <= 3 parts so the recursion ends
    d = np.zeros(n)
    np.add.at(d, i, w)
    np.add.at(d, j, w)
    return 1 + (np.rint(d * 50).astype(np.int64) % 3)


@pytest.mark.parametrize("cluster_fn,size_max,max_itr", [("components", 40, 60), ("components", 25, 6), ("degree", 40, 10000),
                                                         ("degree", 100, 10000)])
def test_clusterbreak_on_device_plans_equals_dense_recursion(evp, cluster_fn, size_max, max_itr):
    # the whole caller loop (sim_fn -> quantile -> threshold -> netcluster -> recurse) from device-resident plans
    # gives the same clusters, labels, filtered sequences and convergence flag as the dense restatement driven by
    # the oracle's similarity matrices
    fn = da.connected_components if cluster_fn == "components" else _weighted_degree_mod3
    want = _dense_clusterbreak(evp, lambda s: port.similarityMH(s, 2, 50, 42), fn, 0.8, size_max, 3, max_itr)
    got = da.clusterbreak(evp, fn, thresh_p=0.8, size_max=size_max, size_min=3, max_itr=max_itr, k=2, n_hash=50, seed=42,
                          verbose=False)
    assert [tuple(r) for r in got["clustered_seq"]] == want["rows"]
    assert got["filtered_seq"] == want["filtered"]
    assert got["convergence"] == want["conv"] and got["calls"] == want["itr"]
    assert want["itr"] > 1  # the recursion was exercised


@pytest.mark.parametrize("cuts", [(0, 16), (0, 5, 16), (0, 1, 2, 15, 16)])
def test_relabelling_by_code_row_shards_equals_full(cuts, monkeypatch):
    # multi-rank form: each rank relabels a share of the packed code rows; here the shares are produced one after the
    # other into the same table, which must then drive the match kernel to the same counts as the unsharded plan
    monkeypatch.setenv("DYNA_MH_PACK16", "1")
    n, n_hash, k = 1500, 31, 3
    seqs = [s.decode() for s in synth_peptides(n)]
    seeds = da.hashfamily_seeds(11, n_hash)
    want = da.mh_match_counts(port.mh_signatures(seqs, k, seeds))
    L = _lib.lib()
    res, off = _lib.flatten(seqs)
    plan = L.dyna_mh_plan_create(n, n_hash, 0, n, 0)
    assert plan, _lib.last_error()
    try:
        _lib.check(L.dyna_mh_plan_upload_sequences(plan, _lib.ptr(res, C.c_uint8), _lib.ptr(off, C.c_int64), k, _lib.ptr(seeds, C.c_uint32), None))
        assert L.dyna_mh_plan_code_rows(plan) == 16 and L.dyna_mh_plan_code_row_bytes(plan) % 512 == 0
        for a, b in zip(cuts[:-1], cuts[1:]):
            _lib.check(L.dyna_mh_plan_run_signatures_shard(plan, a, b, None))
        _lib.check(L.dyna_mh_plan_run_match(plan, None))
        got = np.zeros(len(want), dtype=np.uint16)
        _lib.check(L.dyna_mh_plan_fetch_counts(plan, _lib.ptr(got, C.c_uint16), None))
        assert (got == want).all()
        assert L.dyna_mh_plan_run_signatures_shard(plan, 3, 17, None) != 0  # range check
    finally:
        L.dyna_mh_plan_destroy(plan)


def synth_peptides(n):
    from dynaalign_b200 import synth
    return synth.peptides_clustered(n, children=10)


@pytest.mark.parametrize("pack16", ["0", "1"])
def test_match_counts_several_tile_groups_and_slabs(pack16, monkeypatch):
    # more than 64 tile rows: the grouped (L2-friendly) tile order walks several groups, also from a slab that starts
    # in the middle of a group and of a tile
    monkeypatch.setenv("DYNA_MH_PACK16", pack16)
    rng = np.random.default_rng(21)
    n, n_hash = 17000, 6
    sig = rng.integers(0, 3, size=(n, n_hash), dtype=np.uint32)
    for a, b in [(0, n), (4999, 13001), (16500, n)]:
        want = port.mh_match_counts(sig, a, b)
        got = da.mh_match_counts(sig, a, b)
        assert len(got) == len(want) and (got == want).all(), (a, b)
