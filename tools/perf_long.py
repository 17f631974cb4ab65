import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynaalign_b200 import synth
from dynaalign_b200._lib import check, flatten, lib, ptr
L = lib()
n = 1500
seqs = synth.proteins_uniform(n, mean=900.0, sd=60.0, lo=700, hi=1020)
res, off = flatten(seqs)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = L.dyna_nw_plan_create(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, b"BLOSUM62", 10, 4, 0, n, 0)
check(L.dyna_nw_plan_run(p, st)); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); check(L.dyna_nw_plan_run(p, st)); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1); cells = L.dyna_nw_plan_cells(p)
print("long proteins (700-1020 aa) PACK16=%s: %.1f ms, %d launches, %.0f GCUPS" % (os.environ.get("DYNA_NW_PACK16"), ms, L.dyna_nw_plan_launches(p), cells / ms / 1e6))
L.dyna_nw_plan_destroy(p)
