import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynaalign_b200 import synth
from dynaalign_b200._lib import check, flatten, lib, ptr
L = lib(); n = 20000
seqs = synth.proteins_families(n); res, off = flatten(seqs)
pairs = n * (n + 1) // 2
m = torch.empty(pairs, dtype=torch.int32).pin_memory(); l = torch.empty(pairs, dtype=torch.int32).pin_memory()
for it in range(2):
    t0 = time.perf_counter()
    check(L.dyna_nw_pair_stats(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, b"BLOSUM62", 10, 4, 0, n,
                               C.cast(m.data_ptr(), C.POINTER(C.c_uint32)), C.cast(l.data_ptr(), C.POINTER(C.c_uint32))))
    print("total %.1f ms" % ((time.perf_counter() - t0) * 1e3))
