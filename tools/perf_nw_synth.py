"""Development helper (GPU box): device-timed NW on the first n proteins of the config-5 generator under a few
environment switches.  python tools/perf_nw_synth.py n [ENV=VAL[,ENV=VAL] ...]"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynaalign_b200 import synth
from dynaalign_b200 import _lib as _libmod
if os.environ.get("DYNA_AB_LIB"):  # A/B against another build of the library (development only)
    _libmod.LIB_PATH = os.environ["DYNA_AB_LIB"]
from dynaalign_b200._lib import check, flatten, lib, ptr
L = lib()
kind = "proteins"
if ":" in sys.argv[1]:
    kind, sys.argv[1] = sys.argv[1].split(":")
n = int(sys.argv[1])
if kind == "proteins":
    seqs = synth.proteins_families(20000)[:n]
elif kind == "mix":  # proteome-like length mix: log-normal, median 300 residues, a few sequences beyond 1024
    rng = np.random.default_rng(7)
    lens = np.clip(rng.lognormal(np.log(300.0), 0.5, size=n), 30, 1800).astype(int)
    seqs = [synth.RESIDUES20[rng.integers(0, 20, size=int(L))].tobytes() for L in lens]
else:
    seqs = synth.peptides_uniform(n, length=int(kind[3:] or 16))
res, off = flatten(seqs)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
ref = None
for spec in (sys.argv[2:] or [""]):
    envs = dict(kv.split("=") for kv in spec.split(",") if kv)
    os.environ.update(envs)
    p = L.dyna_nw_plan_create(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, b"BLOSUM62", int(os.environ.get("PERF_GO", "10")), int(os.environ.get("PERF_GE", "4")), 0, n, 0)
    check(L.dyna_nw_plan_run(p, st)); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); check(L.dyna_nw_plan_run(p, st)); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts); cells = L.dyna_nw_plan_cells(p)
    mt = np.zeros(L.dyna_nw_plan_pairs(p), dtype=np.uint32); ln = np.zeros_like(mt)
    check(L.dyna_nw_plan_fetch(p, ptr(mt, C.c_uint32), ptr(ln, C.c_uint32), None))
    sig = (int(mt.astype(np.uint64).sum()), int(ln.astype(np.uint64).sum()), int((mt.astype(np.uint64) * (np.arange(mt.size, dtype=np.uint64) % 1000003)).sum()))
    if ref is None: ref = sig
    print("synth n=%d [%s]: %.2f ms (%s), %d launches, %.0f GCUPS, checksum %s" % (
        n, spec, ms, " ".join("%.1f" % t for t in ts), L.dyna_nw_plan_launches(p), cells / ms / 1e6, "same" if sig == ref else "DIFFERENT"), flush=True)
    L.dyna_nw_plan_destroy(p)
    for k in envs: del os.environ[k]
