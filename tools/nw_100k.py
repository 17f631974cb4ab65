"""BASELINE.json's target case for NW: all 5,000,050,000 pairs (diagonal included) of the 100,000 synthetic 16-mers of
config 4, one GPU.  Times the device path, then checks the result through properties and an oracle sample.
Needs ~40 GB of device memory and (with --fetch) 40 GB of host memory.  python tools/nw_100k.py [--n N] [--fetch]"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynaalign_b200 import _lib, synth  # noqa: E402
from dynaalign_b200._lib import check, flatten, lib, ptr  # noqa: E402
import dynaalign_b200 as da  # noqa: E402
from oracle import port  # noqa: E402  (checker only)

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--fetch", action="store_true")
a = ap.parse_args()

seqs = synth.peptides_uniform(a.n)
n = len(seqs)
L = lib()
res, off = flatten(seqs)
t0 = time.time()
plan = L.dyna_nw_plan_create(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, b"BLOSUM62", 10, 4, 0, n, 0)
assert plan, _lib.last_error()
pairs, cells = L.dyna_nw_plan_pairs(plan), L.dyna_nw_plan_cells(plan)
print("plan create %.2f s  pairs %d cells %.3e" % (time.time() - t0, pairs, cells), flush=True)
assert pairs == n * (n + 1) // 2
import torch  # noqa: E402  (device synchronisation only)
L.dyna_nw_plan_destroy(plan)
t0 = time.time()
plan = L.dyna_nw_plan_create(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, b"BLOSUM62", 10, 4, 0, n, 0)
print("plan create again (allocator warm) %.2f s" % (time.time() - t0), flush=True)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.time()
    check(L.dyna_nw_plan_run(plan, None))
    torch.cuda.synchronize()
    dt = time.time() - t0
    print("run %d: %.3f s  %.1f GCUPS  launches %d" % (rep, dt, cells / dt / 1e9, L.dyna_nw_plan_launches(plan)), flush=True)
idx = lambda i, j: i * n - i * (i - 1) // 2 + (j - i)
rng = np.random.default_rng(5)
if a.fetch:
    mt = np.zeros(pairs, dtype=np.uint32)
    ln = np.zeros(pairs, dtype=np.uint32)
    t0 = time.time()
    check(L.dyna_nw_plan_fetch(plan, ptr(mt, C.c_uint32), ptr(ln, C.c_uint32), None))
    print("fetch %.2f s" % (time.time() - t0), flush=True)
    diag = np.array([idx(i, i) for i in range(n)])
    assert (mt[diag] == 16).all() and (ln[diag] == 16).all()
    ii, jj = rng.integers(0, n, 4000), rng.integers(0, n, 4000)
    i_, j_ = np.minimum(ii, jj), np.maximum(ii, jj)
    for x, y in list(zip(i_.tolist(), j_.tolist())) + [(0, n - 1), (n - 2, n - 1), (n - 1, n - 1), (n // 2, n // 2 + 1)]:
        assert (int(mt[idx(x, y)]), int(ln[idx(x, y)])) == port.nw_pair(seqs[x], seqs[y]), (x, y)
    assert (ln >= 16).all() and (ln <= 32).all() and (mt <= 16).all()
    print("fetched result: diagonal, bounds and 4004 oracle pairs OK", flush=True)
L.dyna_nw_plan_destroy(plan)
# row blocks re-run through the host-buffer entry point: first, one across the 2^32-pair offset, last
for r0, r1 in [(0, 40), (int(n * 0.6), int(n * 0.6) + 40), (n - 300, n)]:
    if r0 < 0:
        continue
    m2, l2 = da.nw_pair_stats(seqs, row_begin=r0, row_end=r1)
    base = idx(r0, r0)
    for _ in range(300):
        x = int(rng.integers(r0, r1))
        y = int(rng.integers(x, n))
        assert (int(m2[idx(x, y) - base]), int(l2[idx(x, y) - base])) == port.nw_pair(seqs[x], seqs[y]), (x, y)
    if a.fetch:
        assert (m2 == mt[base:base + len(m2)]).all() and (l2 == ln[base:base + len(l2)]).all()
    print("rows [%d,%d): %d pairs, oracle sample%s OK" % (r0, r1, len(m2), " and full-run slice" if a.fetch else ""), flush=True)
