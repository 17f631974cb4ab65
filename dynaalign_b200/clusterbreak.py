"""Host-side mirror of the caller loop around the hot path: netcluster + clusterbreak (R/clusterbreak.R:112-136,
:180-275), driven from device-resident MinHash plans instead of dense n x n double matrices.

What stays on the GPU per recursion node: signatures (hashed once for the whole input; a sub-cluster's plan gathers
its rows), the u16 match counts, their histogram and the thresholded edge list.  What comes back to the host: the
(n_hash+1)-bin histogram, the threshold and the edges with similarity >= threshold -- the sparse form of
`pep.sim[pep.sim < threshold] <- 0` (R/clusterbreak.R:219-221).

The community detection itself is the reference's third-party step (igraph::cluster_louvain, randomised, not part of
the path): `cluster_fn` is therefore a required argument here.  It receives the graph as
(n, i, j, weight) -- 0-based vertex ids, one entry per edge, self-loops included exactly as
graph_from_adjacency_matrix(mode = "upper") reads the diagonal of the similarity matrix -- and returns one positive
integer cluster id per vertex.  `connected_components` is a deterministic stand-in used by the tests.
"""
import sys
import time

import numpy as np

from .api import MinHashPlan, NWPlan, RError


def connected_components(n, i, j, weight=None):
    """Deterministic cluster_fn: connected components of the thresholded graph, numbered 1.. in order of first vertex.

    Vectorised (scipy.sparse.csgraph): the stand-in must not dominate the config-3 timing, which is about the similarity
    and threshold steps -- igraph's Louvain is the reference's third-party step and is not installed here."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components as _cc

    i = np.asarray(i, dtype=np.int64)
    j = np.asarray(j, dtype=np.int64)
    graph = coo_matrix((np.ones(i.shape[0], dtype=np.int8), (i, j)), shape=(n, n))
    _, labels = _cc(graph, directed=False)
    _, first, inv = np.unique(labels, return_index=True, return_inverse=True)
    order = np.argsort(np.argsort(first))  # components numbered by their smallest vertex
    return (order[inv] + 1).astype(np.int64)


def netcluster_edges(n, ei, ej, weight, cluster_fn, cluster_wt=True, diag_weight=1.0):
    """netcluster (R/clusterbreak.R:112-136) on an edge list.

    The reference builds graph_from_adjacency_matrix(pepmat, mode = "upper", weighted = TRUE): one edge per non-zero
    entry of the upper triangle INCLUDING the diagonal, so every vertex carries a self-loop of weight sim[i, i]
    (1.0 for similarityMH, src/minHash.cpp:161; `diag_weight` may be a vector: the self-alignment identities of
    similarityNW).  The same graph is handed to `cluster_fn` here.
    """
    loops = np.arange(n, dtype=np.int64)
    gi = np.concatenate([loops, np.asarray(ei, dtype=np.int64)])
    gj = np.concatenate([loops, np.asarray(ej, dtype=np.int64)])
    gw = np.concatenate([np.broadcast_to(np.asarray(diag_weight, dtype=np.float64), (n,)), np.asarray(weight, dtype=np.float64)])
    out = cluster_fn(n, gi, gj, gw) if cluster_wt else cluster_fn(n, gi, gj, None)
    try:
        out = np.asarray(out)
        ok = out.ndim == 1 and out.shape[0] == n and np.issubdtype(out.dtype, np.number)
    except Exception:  # ragged or non-numeric output
        ok = False
    if not ok:
        raise RError("Wrong clustering output format. Output should be a numeric vector of cluster assignment.")
    return out.astype(np.int64)


class _NWNode:
    """A clusterbreak recursion node over ONE computed NW triangle.  similarityNW(sub-cluster) is the sub-matrix of the
    root's similarity matrix for the members in their original order (pair results do not depend on the other sequences,
    and the row sequence of a pair is still the lower-index one), so a node is only a member list: its histogram,
    threshold, edges and self-loop weights are read off the root's slab on the device and nothing is re-aligned -- the
    reference re-runs the whole O(n^2) alignment at every node (R/clusterbreak.R:217, :250-254)."""

    def __init__(self, plan, members=None, owner=True):
        self.plan, self.members, self.owner = plan, members, owner

    def threshold_edges(self, thresh_p):
        return self.plan.threshold_edges(thresh_p, self.members)

    def diag_weight(self):
        m, l = self.plan.diagonal(self.members)
        with np.errstate(divide="ignore", invalid="ignore"):
            return m.astype(np.float64) / l.astype(np.float64)  # NaN for an empty sequence, as src/pairwiseSeqAlign.cpp:311

    def subset(self, local):
        base = np.arange(self.plan.n, dtype=np.int32) if self.members is None else self.members
        return _NWNode(self.plan, np.ascontiguousarray(base[np.asarray(local, dtype=np.int64)], dtype=np.int32), owner=False)

    def close(self):
        if self.owner:
            self.plan.close()


def _log(msg, level="INFO", stream=None):
    print("[%s] %s: %s" % (time.strftime("%H:%M:%S"), level, msg), file=stream or sys.stdout)


def clusterbreak(pep, cluster_fn, thresh_p=0.8, size_max=10, size_min=3, max_itr=10000, k=2, n_hash=50, seed=None,
                 seeds=None, cluster_wt=True, device=0, verbose=True, sim="MH", matrixName="BLOSUM62", gapOpen=10, gapExt=4):
    """clusterbreak (R/clusterbreak.R:180-275) with sim_fn kept on the device: sim = "MH" is similarityMH(k, n_hash)
    (the reference's default sim_fn), sim = "NW" is similarityNW(matrixName, gapOpen, gapExt), aligned ONCE for the
    whole input (every recursion node reads its sub-matrix of that triangle, see _NWNode).

    Same control flow as the reference: one call of `sim_fn` per recursion node, type-7 quantile threshold, clusters
    larger than `size_max` are re-clustered (depth first, in order of first appearance), clusters smaller than
    `size_min` are filtered, labels are "<call number>.<cluster id>".  Differences, both deliberate: the hash seeds are
    drawn once for the whole run instead of once per node (the reference draws a fresh std::random_device seed inside
    every similarityMH call, which makes its recursion irreproducible), and a node that keeps exactly one row no longer
    trips the reference's `nrow(vector)` error.

    Returns dict(clustered_seq = object array [m, 2] of (sequence, label), filtered_seq = list of sequences,
                 convergence = 0/1, calls = number of recursion nodes).
    """
    pep = list(pep)
    if size_max <= size_min:
        raise RError("size_max must be greater than size_min")
    if len(pep) == 0:
        raise RError("empty input sequence vector")
    state = {"rows": [], "itr": 1, "convergence": 1, "filtered": []}
    if sim == "MH":
        root = MinHashPlan(pep, k, n_hash, seed=seed, seeds=seeds, device=device)
    elif sim == "NW":
        root = _NWNode(NWPlan(pep, matrixName, gapOpen, gapExt, device=device).run())
    else:
        raise RError("sim must be 'MH' or 'NW'")

    def recurse(plan, members):
        # `members`: indices into `pep` of this node's sequences, in the node's own order
        if state["itr"] > max_itr:
            if verbose:
                _log("Maximum function calls reached", "WARNING")
            state["convergence"] = 0
            return
        n = len(members)
        if n < 2:  # upper.tri of a 1 x 1 matrix is empty: no threshold, no edges, a single self-loop
            ei = ej = np.zeros(0, dtype=np.int64)
            w = np.zeros(0)
        else:
            _, ei, ej, w = plan.threshold_edges(thresh_p)
        # self-loops: 1.0 for similarityMH (src/minHash.cpp:161), the computed self-alignment identity for similarityNW
        dw = plan.diag_weight() if sim == "NW" else 1.0
        c_index = netcluster_edges(n, ei, ej, w, cluster_fn, cluster_wt, diag_weight=dw)
        c_size = np.bincount(c_index[c_index > 0], minlength=1)[1:]  # tabulate(): ids 1..max
        ids = np.arange(1, len(c_size) + 1)
        id_itr = ids[c_size > size_max]
        id_rm = ids[c_size < size_min]
        in_rm = np.isin(c_index, id_rm)
        in_itr = np.isin(c_index, id_itr)
        state["filtered"].extend(pep[members[t]] for t in np.nonzero(in_rm)[0])
        keep = ~in_rm & ~in_itr
        label = state["itr"]
        state["rows"].extend((pep[members[t]], "%d.%d" % (label, c_index[t])) for t in np.nonzero(keep)[0])
        if len(id_itr) == 0:
            return
        # clusters above size_max, in order of first appearance (unique() on the filtered rows)
        seen = []
        for t in np.nonzero(in_itr)[0]:
            if c_index[t] not in seen:
                seen.append(int(c_index[t]))
        for cid in seen:
            local = np.nonzero(c_index == cid)[0]
            state["itr"] += 1
            child = plan.subset(local)
            try:
                recurse(child, [members[t] for t in local])
            finally:
                child.close()

    try:
        recurse(root, list(range(len(pep))))
    finally:
        root.close()
    if verbose:
        print("\nClustering complete:" if state["convergence"] == 1
              else "\nClustering incomplete, consider adjusting parameters:")
        print("Total function calls (clusters broken): %d" % state["itr"])
    clustered = np.empty((len(state["rows"]), 2), dtype=object)
    for r, row in enumerate(state["rows"]):
        clustered[r, 0], clustered[r, 1] = row
    return {"clustered_seq": clustered, "filtered_seq": list(state["filtered"]), "convergence": state["convergence"],
            "calls": state["itr"]}
