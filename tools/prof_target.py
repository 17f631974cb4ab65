"""Development helper (GPU box): tiny single-shot workloads for ncu captures.  usage: prof_target.py nw|pep|mh [n]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynaalign_b200 import _lib, synth  # noqa: E402
from dynaalign_b200._lib import check, flatten, lib, ptr  # noqa: E402

L = lib()
what = sys.argv[1] if len(sys.argv) > 1 else "nw"
if what in ("nw", "pep"):
    n = int(sys.argv[2]) if len(sys.argv) > 2 else (700 if what == "nw" else 12000)
    seqs = synth.proteins_families(n) if what == "nw" else synth.peptides_uniform(n)
    res, off = flatten(seqs)
    p = L.dyna_nw_plan_create(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, b"BLOSUM62", 10, 4, 0, n, 0)
    assert p, _lib.last_error()
    for _ in range(2):
        check(L.dyna_nw_plan_run(p, None))
    mt = np.zeros(L.dyna_nw_plan_pairs(p), dtype=np.uint32)
    ln = np.zeros_like(mt)
    check(L.dyna_nw_plan_fetch(p, ptr(mt, C.c_uint32), ptr(ln, C.c_uint32), None))
    print("nw", n, L.dyna_nw_plan_cells(p), int(mt.sum()), int(ln.sum()))
    L.dyna_nw_plan_destroy(p)
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    seqs = synth.peptides_uniform(n)
    res, off = flatten(seqs)
    seeds = np.zeros(500, dtype=np.uint32)
    check(L.dyna_hashfamily_seeds(42, 500, ptr(seeds, C.c_uint32)))
    p = L.dyna_mh_plan_create(n, 500, 0, n, 0)
    assert p, _lib.last_error()
    check(L.dyna_mh_plan_upload_sequences(p, ptr(res, C.c_uint8), ptr(off, C.c_int64), 4, ptr(seeds, C.c_uint32), None))
    for _ in range(2):
        check(L.dyna_mh_plan_run_signatures(p, None))
        check(L.dyna_mh_plan_run_match(p, None))
    out = np.zeros(L.dyna_mh_plan_pairs(p), dtype=np.uint16)
    check(L.dyna_mh_plan_fetch_counts(p, ptr(out, C.c_uint16), None))
    print("mh", n, int(out.astype(np.int64).sum()))
    L.dyna_mh_plan_destroy(p)
