"""GPU parity: clusterbreak's threshold + sparsify step (R/clusterbreak.R:217-221) and its recursion
(R/clusterbreak.R:250-254) for sim_fn = similarityNW, read off ONE device-resident NW triangle, against the dense
restatement driven by the oracle's similarity matrices."""
import numpy as np
import pytest

import dynaalign_b200 as da
from conftest import random_seqs
from oracle import port
from oracle.quantile_r import quantile_type7
from test_gpu_minhash import _dense_clusterbreak, _weighted_degree_mod3

pytestmark = pytest.mark.gpu


def family_seqs(rng, parents, children, lo, hi, sub=0.12):
    """parents x children sequences with substitutions and a few indels, shuffled: identities spread over (0, 1]."""
    al = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)
    out = []
    for _ in range(parents):
        p = al[rng.integers(0, 20, size=int(rng.integers(lo, hi + 1)))]
        for _ in range(children):
            c = p.copy()
            hit = rng.random(len(c)) < sub
            c[hit] = al[rng.integers(0, 20, size=int(hit.sum()))]
            if rng.random() < 0.5 and len(c) > 4:
                cut = int(rng.integers(1, len(c) - 1))
                c = np.delete(c, cut)
            out.append(c.tobytes().decode())
    order = rng.permutation(len(out))
    return [out[t] for t in order]


def dense_edges(sim, p):
    n = sim.shape[0]
    thr = quantile_type7(sim[np.triu_indices(n, 1)], p)
    dense = sim.copy()
    dense[dense < thr] = 0.0  # pep.sim[pep.sim < threshold] <- 0
    wi, wj = np.nonzero(np.triu(dense, 1))
    return thr, wi, wj, dense[wi, wj]


@pytest.mark.parametrize("p", [0.0, 0.31, 0.8, 0.9371, 1.0])
def test_threshold_and_edge_list_nw(p):
    rng = np.random.default_rng(11)
    seqs = family_seqs(rng, 12, 9, 18, 70)
    full = port.similarityNW(seqs)
    want_thr, wi, wj, ww = dense_edges(full, p)
    thr, ei, ej, w = da.similarityNW_edges(seqs, thresh_p=p)
    assert thr == want_thr
    assert (ei == wi).all() and (ej == wj).all() and (w == ww).all()


def test_histogram_is_the_triangle_and_ranks_add():
    rng = np.random.default_rng(12)
    seqs = random_seqs(rng, 90, 1, 40) + ["", "A"]
    n = len(seqs)
    wm, wl = port.nw_pair_stats(seqs)
    iu = np.triu_indices(n)  # packed triangle incl. diagonal, row-major
    strict = iu[0] < iu[1]
    ml = max(len(s) for s in seqs)
    want = np.zeros((ml + 1, 2 * ml + 1), dtype=np.uint64)
    np.add.at(want, (wm[strict], wl[strict]), 1)
    plan = da.NWPlan(seqs).run()
    assert plan.max_len == ml
    assert (plan.stat_histogram() == want).all()
    dm, dl = plan.diagonal()
    assert (dm == wm[~strict]).all() and (dl == wl[~strict]).all()
    plan.close()
    # row slabs: histograms, edge lists and diagonals of the slabs add up to the whole
    b = da.partition_rows(n, 3, [len(s) for s in seqs])
    thr = da.quantile_type7_identities(want, 0.6)
    full = port.similarityNW(seqs)
    assert thr == quantile_type7(full[np.triu_indices(n, 1)], 0.6)
    hs, es, ds = np.zeros_like(want), [], np.zeros(n, dtype=np.uint32)
    for s in range(3):
        part = da.NWPlan(seqs, row_begin=int(b[s]), row_end=int(b[s + 1])).run()
        hs += part.stat_histogram()
        es.append(part.edges_at(thr, 10 ** 5))
        ds += part.diagonal()[1]
        part.close()
    assert (hs == want).all() and (ds == wl[~strict]).all()
    dense = full.copy()
    dense[dense < thr] = 0.0
    wi, wj = np.nonzero(np.triu(np.nan_to_num(dense), 1))
    assert (np.concatenate([e[0] for e in es]) == wi).all() and (np.concatenate([e[1] for e in es]) == wj).all()


def test_two_empty_sequences_make_the_quantile_fail_like_r():
    plan = da.NWPlan(["", "ARND", ""]).run()
    with pytest.raises(da.DynaAlignError, match="missing values and NaN's not allowed"):
        plan.threshold_edges(0.5)
    plan.close()


def test_node_subsets_are_submatrices():
    # clusterbreak's recursion: similarityNW(sub-cluster) == the sub-matrix of the root's matrix (members in order)
    rng = np.random.default_rng(13)
    seqs = family_seqs(rng, 10, 12, 20, 90)
    plan = da.NWPlan(seqs).run()
    for size in (2, 17, 75):
        idx = np.sort(rng.choice(len(seqs), size=size, replace=False))
        sub = port.similarityNW([seqs[t] for t in idx])  # what the reference computes again for the node
        for p in (0.5, 0.8):
            want_thr, wi, wj, ww = dense_edges(sub, p)
            thr, ei, ej, w = plan.threshold_edges(p, idx)
            assert thr == want_thr
            assert (ei == wi).all() and (ej == wj).all() and (w == ww).all()
        dm, dl = plan.diagonal(idx)
        assert (dm / dl == np.diag(sub)).all()
    with pytest.raises(da.DynaAlignError, match="strictly increasing"):
        plan.stat_histogram(np.array([3, 3, 5]))
    with pytest.raises(da.DynaAlignError, match="strictly increasing"):
        plan.stat_histogram(np.array([0, len(seqs)]))
    plan.close()


@pytest.mark.parametrize("cluster_fn,size_max,max_itr", [("components", 30, 60), ("components", 20, 4), ("degree", 40, 10000)])
def test_clusterbreak_nw_on_one_triangle_equals_dense_recursion(cluster_fn, size_max, max_itr):
    # the whole caller loop with sim_fn = similarityNW: the reference re-aligns every recursion node; here every node is
    # read off the root's triangle -- same clusters, labels, filtered sequences, convergence flag and call count
    rng = np.random.default_rng(14)
    seqs = family_seqs(rng, 9, 14, 15, 60)
    fn = da.connected_components if cluster_fn == "components" else _weighted_degree_mod3
    want = _dense_clusterbreak(seqs, lambda s: port.similarityNW(s), fn, 0.8, size_max, 3, max_itr)
    got = da.clusterbreak(seqs, fn, thresh_p=0.8, size_max=size_max, size_min=3, max_itr=max_itr, sim="NW", verbose=False)
    assert [tuple(r) for r in got["clustered_seq"]] == want["rows"]
    assert got["filtered_seq"] == want["filtered"]
    assert got["convergence"] == want["conv"] and got["calls"] == want["itr"]
    assert want["itr"] > 1  # the recursion was exercised
