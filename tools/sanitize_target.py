"""Small run through every kernel class, for compute-sanitizer (one tool per gpurun call):

    compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/sanitize_target.py

Covers the places where shared memory is reused across proxies or warps: the TMA match kernel with several tiles per
CTA (its staging tile aliases the TMA ring across the generic->async proxy fence), the multi-pass NW kernels (in-place
scratch lines), the cooperative two-warp NW kernel (shared-memory ring with release/acquire counters), the tiled
expansion (shared-memory transpose), plus every other kernel once."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dynaalign_b200 as da  # noqa: E402

rng = np.random.default_rng(0)
al = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)
mk = lambda L: al[rng.integers(0, 20, size=int(L))].tobytes().decode()
# NW: empty row, thread2 (<=32), warp2 (33..384), cooperative (385..768), packed multi-pass (> 768 rows, or columns > 1024),
#     32-bit warp + multi-pass (PACK16=0), unslanted kernels
seqs = [mk(L) for L in (0, 5, 12, 31, 40, 333, 340, 566, 566, 570, 700, 900, 1100)]
a = da.similarityNW(seqs)
os.environ["DYNA_NW_PACK16"] = "0"
b = da.similarityNW(seqs)  # 32-bit thread / warp / multi-pass kernels
del os.environ["DYNA_NW_PACK16"]
assert a.tobytes() == b.tobytes()
os.environ["DYNA_NW_CO"] = "0"
b = da.similarityNW(seqs)  # tall single-pass strips + packed multi-pass instead of the cooperative kernel
del os.environ["DYNA_NW_CO"]
assert a.tobytes() == b.tobytes()
ha = [mk(566) for _ in range(140)]  # two units of the cooperative kernel (128 + 12 columns), all groups busy
m1, l1 = da.nw_pair_stats(ha[:1] + ha)
ha_stats = (m1, l1)
c = da.similarityNW(seqs[:6], "BLOSUM62", 5, 200)  # unslanted kernels
m8, l8 = da.nw_pair_stats8([mk(16) for _ in range(300)])
# MinHash: u32 path, 16-bit path, several tiles per CTA on both, linear signatures, expansion, histogram + edges, narrow fetch
peps = [mk(16) for _ in range(300)] + ["", "AC"]
mh1 = da.similarityMH(peps, 4, 37, seed=1)
os.environ["DYNA_MH_PACK16"] = "1"
m2 = da.similarityMH(peps, 4, 37, seed=1)
thr, ei, ej, w = da.similarityMH_edges(peps, 2, 20, 0.9, seed=1)
assert mh1.tobytes() == m2.tobytes()
many = [mk(16) for _ in range(6000)]  # 47 x 47 / 2 = 1128 tiles of 128 x 128 pairs: several tiles per resident CTA
p16 = da.MinHashPlan(many, 4, 32, seed=3)
c16 = p16.match_counts()
c8, xi, xc = da.MinHashPlan(many, 4, 32, seed=3).match_counts8()
del os.environ["DYNA_MH_PACK16"]
os.environ["DYNA_MH_PACK16"] = "0"
c32 = da.MinHashPlan(many, 4, 32, seed=3).match_counts()
del os.environ["DYNA_MH_PACK16"]
os.environ["DYNA_MH_MATCH"] = "ldg"   # the match kernel without TMA (no staging tile aliasing the ring): same counts
cl = da.MinHashPlan(many, 4, 32, seed=3).match_counts()
del os.environ["DYNA_MH_MATCH"]
assert (c16 == c32).all() and (c16 == cl).all() and (c8 == np.minimum(c16, 255)).all()
# races show up as run-to-run differences: the kernels with cross-warp or cross-proxy shared memory, ten times each
for _ in range(10):
    assert (da.MinHashPlan(many, 4, 32, seed=3).match_counts() == c32).all()
    mr, lr = da.nw_pair_stats(ha[:1] + ha)
    assert (mr == ha_stats[0]).all() and (lr == ha_stats[1]).all()
r = da.minhash(peps[:40], 3, 16, rng=np.random.default_rng(1))
long_seq = [mk(700) for _ in range(6)]
s = da.mh_signatures(long_seq, 5, da.hashfamily_seeds(3, 12))  # warp-min signature kernel
print("sanitize target ok", a.shape, mh1.shape, len(ei), r["dist_matrix"].shape, s.shape, c16.shape)
