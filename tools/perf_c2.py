"""Development helper (GPU box): device-timed BASELINE config 2 (similarityNW on h3n2sample[1:1000]) under a few
kernel-selection switches.  python tools/perf_c2.py [ENV=VAL[,ENV=VAL] ...]   (each argument is one run)"""
import ctypes as C, gzip, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynaalign_b200 import _lib as _libmod
if os.environ.get("DYNA_AB_LIB"):  # A/B against another build of the library (development only)
    _libmod.LIB_PATH = os.environ["DYNA_AB_LIB"]
from dynaalign_b200._lib import check, flatten, lib, ptr
L = lib()
with gzip.open(os.path.join(ROOT, "tests/golden/h3n2sample_first1000.json.gz"), "rt") as f:
    d = json.load(f)
h3 = [d["unique"][i] for i in d["index"]]
res, off = flatten(h3); n = len(h3)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
gold = np.load(os.path.join(ROOT, "tests/golden/nw_h3n2_1000_stats.npz"))
runs = sys.argv[1:] or ["", "DYNA_NW_CO=0"]
for spec in runs:
    envs = dict(kv.split("=") for kv in spec.split(",") if kv)
    os.environ.update(envs)
    p = L.dyna_nw_plan_create(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, b"BLOSUM62", 10, 4, 0, n, 0)
    check(L.dyna_nw_plan_run(p, st)); torch.cuda.synchronize()
    ts = []
    for _ in range(int(os.environ.get('PERF_C2_REPS', '5'))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); check(L.dyna_nw_plan_run(p, st)); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts); cells = L.dyna_nw_plan_cells(p)
    mt = np.zeros(L.dyna_nw_plan_pairs(p), dtype=np.uint32); ln = np.zeros_like(mt)
    check(L.dyna_nw_plan_fetch(p, ptr(mt, C.c_uint32), ptr(ln, C.c_uint32), None))
    ok = bool((mt == gold["matches"]).all() and (ln == gold["length"]).all())
    print("config2 [%s]: %.2f ms (%s), %d launches, %.0f GCUPS, golden %s" % (
        spec, ms, " ".join("%.2f" % t for t in ts), L.dyna_nw_plan_launches(p), cells / ms / 1e6, "OK" if ok else "MISMATCH"), flush=True)
    L.dyna_nw_plan_destroy(p)
    for k in envs: del os.environ[k]
