"""Host logic of the clusterbreak mirror (R/clusterbreak.R:112-136,180-275) that needs no GPU."""
import numpy as np
import pytest

import dynaalign_b200 as da
from dynaalign_b200.api import RError


def test_connected_components_numbering():
    # components are numbered by their smallest vertex, independent of edge order
    i = np.array([4, 1, 2])
    j = np.array([5, 3, 0])
    assert da.connected_components(6, i, j).tolist() == [1, 2, 1, 2, 3, 3]
    assert da.connected_components(6, i[::-1], j[::-1]).tolist() == [1, 2, 1, 2, 3, 3]
    assert da.connected_components(3, [], []).tolist() == [1, 2, 3]


def test_netcluster_graph_has_the_diagonal_self_loops():
    # graph_from_adjacency_matrix(mode = "upper") reads the diagonal too (R/clusterbreak.R:122-124)
    seen = {}

    def fn(n, i, j, w):
        seen.update(n=n, i=i.tolist(), j=j.tolist(), w=None if w is None else w.tolist())
        return [1] * n

    out = da.netcluster_edges(3, [0], [2], [0.5], fn)
    assert out.tolist() == [1, 1, 1]
    assert seen == {"n": 3, "i": [0, 1, 2, 0], "j": [0, 1, 2, 2], "w": [1.0, 1.0, 1.0, 0.5]}
    da.netcluster_edges(3, [0], [2], [0.5], fn, cluster_wt=False)
    assert seen["w"] is None


@pytest.mark.parametrize("bad", [lambda n, i, j, w: "abc", lambda n, i, j, w: [1] * (n + 1), lambda n, i, j, w: [[1] * n]])
def test_netcluster_rejects_malformed_membership(bad):
    with pytest.raises(RError, match="Wrong clustering output format"):
        da.netcluster_edges(3, [], [], [], bad)


def test_clusterbreak_argument_errors():
    with pytest.raises(RError, match="size_max must be greater than size_min"):
        da.clusterbreak(["AAAA"], da.connected_components, size_max=3, size_min=3)
    with pytest.raises(RError, match="empty input sequence vector"):
        da.clusterbreak([], da.connected_components)
