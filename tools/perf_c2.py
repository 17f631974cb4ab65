import ctypes as C, gzip, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynaalign_b200._lib import check, flatten, lib, ptr
L = lib()
with gzip.open(os.path.join(ROOT, "tests/golden/h3n2sample_first1000.json.gz"), "rt") as f:
    d = json.load(f)
h3 = [d["unique"][i] for i in d["index"]]
res, off = flatten(h3); n = len(h3)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for env in (None, "0"):
    if env is not None: os.environ["DYNA_NW_PACK16"] = env
    p = L.dyna_nw_plan_create(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, b"BLOSUM62", 10, 4, 0, n, 0)
    check(L.dyna_nw_plan_run(p, st)); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); check(L.dyna_nw_plan_run(p, st)); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); cells = L.dyna_nw_plan_cells(p)
    print("config2 PACK16=%s: %.1f ms, %d launches, %.0f GCUPS" % (env, ms, L.dyna_nw_plan_launches(p), cells / ms / 1e6))
    L.dyna_nw_plan_destroy(p)
