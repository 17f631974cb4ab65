"""Tiny run through every kernel class, for compute-sanitizer (one tool per gpurun call)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dynaalign_b200 as da  # noqa: E402

rng = np.random.default_rng(0)
al = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)
mk = lambda L: al[rng.integers(0, 20, size=int(L))].tobytes().decode()
# NW: empty row, thread2 (<=32), warp2 (33..640), 32-bit warp (641..768), multipass (>768)
seqs = [mk(L) for L in (0, 5, 12, 31, 40, 333, 700, 900)]
a = da.similarityNW(seqs)
os.environ["DYNA_NW_PACK16"] = "0"
b = da.similarityNW(seqs)  # 32-bit thread / warp kernels
del os.environ["DYNA_NW_PACK16"]
assert a.tobytes() == b.tobytes()
c = da.similarityNW(seqs[:6], "BLOSUM62", 5, 200)  # unslanted kernels
# MinHash: u32 path, 16-bit path, linear signatures, expansion, histogram + edges
peps = [mk(16) for _ in range(300)] + ["", "AC"]
m1 = da.similarityMH(peps, 4, 37, seed=1)
os.environ["DYNA_MH_PACK16"] = "1"
m2 = da.similarityMH(peps, 4, 37, seed=1)
thr, ei, ej, w = da.similarityMH_edges(peps, 2, 20, 0.9, seed=1)
del os.environ["DYNA_MH_PACK16"]
assert m1.tobytes() == m2.tobytes()
r = da.minhash(peps[:40], 3, 16, rng=np.random.default_rng(1))
long_seq = [mk(700) for _ in range(6)]
s = da.mh_signatures(long_seq, 5, da.hashfamily_seeds(3, 12))  # warp-min signature kernel
print("sanitize target ok", a.shape, m1.shape, len(ei), r["dist_matrix"].shape, s.shape)
