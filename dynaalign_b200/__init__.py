"""dynaalign_b200: B200-native (sm_100a) all-pairs similarity hot path of DynaAlign.

``similarityNW`` / ``similarityMH`` / ``minhash`` with the reference's interface, backed by hand-written CUDA
kernels behind a C ABI (include/dynaalign_b200.h, dynaalign_b200/csrc).  No CPU fallback.
"""
from ._lib import DynaAlignError, LIB_PATH  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import __all__ as _api_all
from .clusterbreak import clusterbreak, connected_components, netcluster_edges  # noqa: F401

__all__ = list(_api_all) + ["LIB_PATH", "clusterbreak", "connected_components", "netcluster_edges"]
__version__ = "0.1.0"
