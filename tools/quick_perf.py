"""Development helper (GPU box): integer-issue probe and quick device-timed throughput of both kernels."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynaalign_b200 import _lib, synth  # noqa: E402
from dynaalign_b200._lib import check, flatten, lib, ptr  # noqa: E402

L = lib()
torch.cuda.init()
st = torch.cuda.current_stream().cuda_stream


def probe():
    names = ["IADD3", "VIADDMNMX", "VIMNMX3", "NW-mix(5)", "ISETP+IADD(3)"]
    for k in range(5):
        ops, ms = C.c_double(0), C.c_double(0)
        check(L.dyna_probe_int_issue(k, C.byref(ops), C.byref(ms), C.c_void_p(st)))
        print("probe %-14s %8.2f T lane-ops/s  (%.2f ms)" % (names[k], ops.value / 1e12, ms.value), flush=True)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), ts


def mh(n, n_hash=500, k=4, mode=None):
    if mode:
        os.environ["DYNA_MH_MATCH"] = mode
    seqs = synth.peptides_uniform(n)
    res, off = flatten(seqs)
    seeds = np.zeros(n_hash, dtype=np.uint32)
    check(L.dyna_hashfamily_seeds(42, n_hash, ptr(seeds, C.c_uint32)))
    p = L.dyna_mh_plan_create(n, n_hash, 0, n, 0)
    assert p, _lib.last_error()
    check(L.dyna_mh_plan_upload_sequences(p, ptr(res, C.c_uint8), ptr(off, C.c_int64), k, ptr(seeds, C.c_uint32), C.c_void_p(st)))
    t_sig, _ = timed(lambda: check(L.dyna_mh_plan_run_signatures(p, C.c_void_p(st))))
    t_m, ts = timed(lambda: check(L.dyna_mh_plan_run_match(p, C.c_void_p(st))))
    pairs = L.dyna_mh_plan_pairs(p)
    print("MH n=%d n_hash=%d mode=%s: signatures %.3f ms, match %.3f ms %s -> %.3e pairs/s, %.2f Tcmp/s, %.1f GB/s(alg)" % (
        n, n_hash, mode or "tma", t_sig, t_m, ["%.2f" % t for t in ts], pairs / t_m * 1e3, pairs * n_hash / t_m * 1e3 / 1e12,
        (2.0 * pairs + 4.0 * n * n_hash) / t_m * 1e3 / 1e9), flush=True)
    L.dyna_mh_plan_destroy(p)


def nw(n, kind="families"):
    if kind == "pep12":
        seqs = synth.peptides_uniform(n, length=12)
    elif kind == "pep16":
        seqs = synth.peptides_uniform(n, length=16)
    else:
        seqs = synth.proteins_families(n) if kind == "families" else synth.proteins_uniform(n)
    res, off = flatten(seqs)
    t0 = time.time()
    p = L.dyna_nw_plan_create(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, b"BLOSUM62", 10, 4, 0, n, 0)
    assert p, _lib.last_error()
    t_plan = time.time() - t0
    t, ts = timed(lambda: check(L.dyna_nw_plan_run(p, C.c_void_p(st))), reps=2)
    cells = L.dyna_nw_plan_cells(p)
    print("NW n=%d (%s): plan %.2f s, run %.1f ms %s, %d launches -> %.1f GCUPS" % (
        n, kind, t_plan, t, ["%.1f" % x for x in ts], L.dyna_nw_plan_launches(p), cells / t * 1e3 / 1e9), flush=True)
    L.dyna_nw_plan_destroy(p)


if __name__ == "__main__":
    what = sys.argv[1:] or ["probe", "mh", "nw"]
    print(torch.cuda.get_device_name(0), flush=True)
    if "probe" in what:
        probe()
    if "mh" in what:
        mh(8192)
        mh(8192, mode="ldg")
        mh(32768, mode="tma")
    if "mhfull" in what:
        mh(100000, mode="tma")
    if "nw" in what:
        nw(1000)
        nw(2000, "uniform")
    if "nwbig" in what:
        nw(5000)
    if "nwpep" in what:
        nw(20000, "pep12")
        nw(20000, "pep16")
