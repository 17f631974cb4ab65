#!/usr/bin/env python
"""bench.py -- headline benchmark of the DynaAlign all-pairs similarity hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): NW all-pairs GCUPS and MinHash pairs/s.  One JSON line on stdout (rank 0):

  * headline `metric` = nw_allpairs_gcups on BASELINE config 5 (synthetic 20k proteins of ~330 aa, BLOSUM62 10/4,
    all pairs i<=j, row blocks balanced by DP cells over the ranks);  a step = one pass over the whole workload.
  * `minhash` sub-object = minhash_pairs_per_sec on config 4 (synthetic 100k peptides of 16 aa, k=4, n_hash=500),
    with its own value / e2e / roofline.
  * `value`: inputs already resident in HBM (device plans), CUDA events on the launching stream, max over ranks.
  * `e2e`: the same work through the host C ABI (dyna_nw_pair_stats / plan upload+run+fetch): pinned host buffers,
    H2D + D2H inside the timed region.
  * `roofline`: for the dominant kernel (nw_warp_kernel): algorithmic integer ops (11 per DP cell)
    per second against the INT32 issue peak measured live with dyna_probe_int_issue (MEASURED_PEAKS.json has no
    integer figure).  The MinHash match kernel reports the HBM roofline BASELINE.json names (2.04 B/pair) and the
    integer one that actually binds it.
  * `cpu_baseline` (N=1, rank 0): the reference's own C++ (oracle/_ref, compiled unmodified) timed on a bounded
    sample of the same workloads on the box's host cores.

`--impl reference` times that CPU reference alone on the same metric (rank 0 only under torchrun).
Scaling is STRONG: the workload is fixed and sharded, so per-GPU work shrinks as N grows.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NW_OPS_PER_CELL = 11.0        # algorithmic integer instructions per DP cell
MH_BYTES_PER_PAIR = 2.0       # one u16 match count per unordered pair (+ 4*N*n_hash signature bytes, added below)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            busy = [c for c, p in zip(sm, power) if p > 0.5 * max(power)] or sm
            out.update(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=float(max(power)))
        return out


# --------------------------------------------------------------------------------------------- workloads
def nw_workload(n):
    from dynaalign_b200 import synth
    return synth.proteins_families(n)


def mh_workload(n):
    from dynaalign_b200 import synth
    return synth.peptides_uniform(n)


# --------------------------------------------------------------------------------------------- reference (CPU) arm
def reference_backend():
    """(module, kind): the compiled reference when it travelled with the repo, else the C port of it."""
    from oracle import port, ref
    if ref.available():
        return ref, "reference"
    return port, "port"


def cpu_nw_sample(seqs, n_sample):
    """Reference similarityNW (single-threaded by construction) on the first n_sample sequences -> (GCUPS, seconds)."""
    mod, kind = reference_backend()
    sub = seqs[:n_sample]
    lens = np.array([len(s) for s in sub], dtype=np.int64)
    suffix = np.cumsum(lens[::-1])[::-1]
    cells = int((lens * suffix).sum())
    t0 = time.perf_counter()
    mod.similarityNW(sub, "BLOSUM62", 10, 4)
    dt = time.perf_counter() - t0
    return cells / dt / 1e9, dt, cells, kind


def cpu_mh_sample(seqs, n_sample, threads):
    mod, kind = reference_backend()
    sub = seqs[:n_sample]
    if kind == "reference":
        mod.set_threads(threads)
    else:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    t0 = time.perf_counter()
    mod.similarityMH(sub, 4, 500, 42)
    dt = time.perf_counter() - t0
    pairs = n_sample * (n_sample - 1) // 2
    return pairs / dt, dt, pairs, kind


def run_reference_arm(args, rank):
    if rank != 0:
        return
    seqs = nw_workload(min(args.nw_n, 256))
    n_sample = 64
    times = []
    cells = kind = None
    for it in range(args.warmup + args.steps):
        gc, dt, cells, kind = cpu_nw_sample(seqs, n_sample)
        if it >= args.warmup:
            times.append(dt)
    dt = float(np.mean(times))
    value = cells / dt / 1e9
    peps = mh_workload(8000)
    cores = os.cpu_count() or 1
    mh_rate, mh_dt, mh_pairs, _ = cpu_mh_sample(peps, 8000, cores)
    line = {
        "impl": "reference", "metric": "nw_allpairs_gcups", "value": value, "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "similarityNW BLOSUM62 10/4, synthetic %d proteins ~330 aa (config 5); reference timed on the "
                               "first %d sequences per step" % (args.nw_n, n_sample)},
        "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": 1, "kind": kind,
                         "sample": "first %d sequences of config 5 (%d pairs, %.3g cells) per step; similarityNW is "
                                   "single-threaded in the reference" % (n_sample, n_sample * (n_sample + 1) // 2, cells)},
        "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "minhash": {"metric": "minhash_pairs_per_sec", "value": mh_rate, "unit": "pairs/s", "cores": cores,
                    "sample": "first 8000 peptides of config 4 (%d pairs), k=4 n_hash=500, signature build included" % mh_pairs},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--nw-n", type=int, default=20000, help="config 5 size (development: smaller)")
    ap.add_argument("--mh-n", type=int, default=100000, help="config 4 size (development: smaller)")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist

    from dynaalign_b200 import _lib
    from dynaalign_b200._lib import check, flatten, lib, ptr

    L = lib()
    if L.dyna_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL announces its version on stdout when the communicator is created; stdout carries exactly one JSON line,
        # so route fd 1 to stderr while the process group comes up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    dev = local_rank
    stream = torch.cuda.current_stream()
    st = C.c_void_p(stream.cuda_stream)
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    phys = vis.split(",")[local_rank] if vis else str(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def timed_steps(fn, warmup, steps):
        """K steps, each bracketed by events on the launching stream; L2 flushed between steps (outside the events)."""
        for _ in range(warmup):
            fn()
        barrier()
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for k in range(steps):
            flush_buf.fill_(k)
            e0[k].record(stream)
            fn()
            e1[k].record(stream)
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in zip(e0, e1))
        return max_over_ranks(ms)

    # ---------------- integer issue peak (roofline denominator), measured live on this device
    peak_ops, peak_ms = C.c_double(0), C.c_double(0)
    check(L.dyna_probe_int_issue(0, C.byref(peak_ops), C.byref(peak_ms), st))
    int_peak = peak_ops.value  # lane-ops/s

    sampler = ClockSampler(phys)
    launches = 0

    # ================================================================== NW, config 5
    seqs = nw_workload(args.nw_n)
    n = len(seqs)
    res, off = flatten(seqs)
    lens = np.diff(off)
    bounds = np.zeros(world + 1, dtype=np.int64)
    check(L.dyna_partition_rows(n, ptr(lens.astype(np.int64), C.c_int64), 1, world, ptr(bounds, C.c_int64)))
    rb, re_ = int(bounds[rank]), int(bounds[rank + 1])
    plan = L.dyna_nw_plan_create(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, b"BLOSUM62", 10, 4, rb, re_, dev)
    if not plan:
        raise SystemExit("bench.py: " + _lib.last_error())
    my_cells = L.dyna_nw_plan_cells(plan)
    my_pairs = L.dyna_nw_plan_pairs(plan)
    total_cells = sum_over_ranks(my_cells)
    total_pairs = sum_over_ranks(my_pairs)

    sampler.start()
    nw_ms = timed_steps(lambda: check(L.dyna_nw_plan_run(plan, st)), args.warmup, args.steps)
    clocks = sampler.stop()
    nw_launch_per_step = L.dyna_nw_plan_launches(plan)
    launches += nw_launch_per_step * args.steps
    nw_ms_per_step = nw_ms / args.steps
    nw_gcups = total_cells / (nw_ms_per_step * 1e-3) / 1e9
    L.dyna_nw_plan_destroy(plan)

    # e2e: host C ABI with pinned host buffers, H2D + D2H inside the timed region
    pin_res = torch.from_numpy(res).pin_memory()
    pin_off = torch.from_numpy(off).pin_memory()
    out_m = torch.empty(max(my_pairs, 1), dtype=torch.int32).pin_memory()
    out_l = torch.empty(max(my_pairs, 1), dtype=torch.int32).pin_memory()
    check(L.dyna_set_device(dev))

    def nw_e2e_step():
        check(L.dyna_nw_pair_stats(C.cast(pin_res.data_ptr(), C.POINTER(C.c_uint8)), C.cast(pin_off.data_ptr(), C.POINTER(C.c_int64)),
                                   n, b"BLOSUM62", 10, 4, rb, re_, C.cast(out_m.data_ptr(), C.POINTER(C.c_uint32)),
                                   C.cast(out_l.data_ptr(), C.POINTER(C.c_uint32))))

    nw_e2e_step()  # warm-up (allocator, first touch)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 2))
    for _ in range(e2e_steps):
        nw_e2e_step()
    barrier()
    nw_e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    nw_e2e_gcups = total_cells / nw_e2e_s / 1e9
    launches += nw_launch_per_step * (e2e_steps + 1)
    # sanity: self-alignment of the first row of this rank's slab is an identity (matches == length == len)
    if my_pairs > 0:
        assert int(out_m[0]) == int(lens[rb]) and int(out_l[0]) == int(lens[rb]), "NW e2e sanity check failed"
    nw_h2d = int(res.nbytes + off.nbytes)
    nw_d2h = int(8 * my_pairs)
    del out_m, out_l, pin_res, pin_off

    # ================================================================== MinHash, config 4
    peps = mh_workload(args.mh_n)
    mn, n_hash, k = len(peps), 500, 4
    mres, moff = flatten(peps)
    seeds = np.zeros(n_hash, dtype=np.uint32)
    check(L.dyna_hashfamily_seeds(42, n_hash, ptr(seeds, C.c_uint32)))
    mb = np.zeros(world + 1, dtype=np.int64)
    check(L.dyna_partition_rows(mn, None, 0, world, ptr(mb, C.c_int64)))
    mrb, mre = int(mb[rank]), int(mb[rank + 1])
    mplan = L.dyna_mh_plan_create(mn, n_hash, mrb, mre, dev)
    if not mplan:
        raise SystemExit("bench.py: " + _lib.last_error())
    mh_my_pairs = L.dyna_mh_plan_pairs(mplan)
    mh_total_pairs = sum_over_ranks(mh_my_pairs)
    check(L.dyna_mh_plan_upload_sequences(mplan, ptr(mres, C.c_uint8), ptr(moff, C.c_int64), k, ptr(seeds, C.c_uint32), st))

    # N > 1: the relabelling of the signature rows is sharded across ranks and completed by one NCCL all-gather of the
    # code table (+ a max-reduce of the overflow gate); N = 1 (or DYNA_MH_SHARD_RELABEL=0): plain run_signatures
    from dynaalign_b200.multirank import ShardedSignatures
    msig = ShardedSignatures(mplan, world if os.environ.get("DYNA_MH_SHARD_RELABEL", "1") != "0" else 1, rank, dist, torch,
                             torch.device("cuda", dev))

    def mh_step():
        msig.run(st)
        check(L.dyna_mh_plan_run_match(mplan, st))

    mh_ms = timed_steps(mh_step, max(args.warmup, 3), max(args.steps, 3)) / max(args.steps, 3)
    launches += 3 * max(args.steps, 3)
    # match kernel alone (the roofline kernel of this half)
    msig.run(st)
    mh_match_ms = timed_steps(lambda: check(L.dyna_mh_plan_run_match(mplan, st)), 1, 3) / 3
    launches += 3
    mh_pairs_s = mh_total_pairs / (mh_ms * 1e-3)

    # e2e: upload sequences, signatures, match, counts slab back to pinned host memory
    pin_counts = torch.empty(max(mh_my_pairs, 1), dtype=torch.int16).pin_memory()
    pin_mres = torch.from_numpy(mres).pin_memory()
    pin_moff = torch.from_numpy(moff).pin_memory()

    def mh_e2e_step():
        check(L.dyna_mh_plan_upload_sequences(mplan, C.cast(pin_mres.data_ptr(), C.POINTER(C.c_uint8)),
                                              C.cast(pin_moff.data_ptr(), C.POINTER(C.c_int64)), k, ptr(seeds, C.c_uint32), st))
        msig.run(st)
        check(L.dyna_mh_plan_run_match_fetch(mplan, C.cast(pin_counts.data_ptr(), C.POINTER(C.c_uint16)), st))

    mh_e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        mh_e2e_step()
    barrier()
    mh_e2e_s = max_over_ranks((time.perf_counter() - t0) / 3)
    launches += 3 * (3 + 16 * 2)

    # sparse e2e: what clusterbreak consumes at this size (the dense matrix would be 80 GB) -- the type-7 quantile threshold and
    # the edge list above it (R/clusterbreak.R:219-221).  Host buffers in, histogram + (i, j, count) edges out.
    hist = np.zeros(n_hash + 1, dtype=np.uint64)
    edge_cap = [0]
    pin_ei = pin_ej = pin_ec = None

    def mh_sparse_step(thresh_p=0.8):
        nonlocal pin_ei, pin_ej, pin_ec
        check(L.dyna_mh_plan_upload_sequences(mplan, C.cast(pin_mres.data_ptr(), C.POINTER(C.c_uint8)),
                                              C.cast(pin_moff.data_ptr(), C.POINTER(C.c_int64)), k, ptr(seeds, C.c_uint32), st))
        msig.run(st)
        check(L.dyna_mh_plan_run_match(mplan, st))
        check(L.dyna_mh_plan_count_histogram(mplan, ptr(hist, C.c_uint64), st))
        ghist = hist
        if world > 1:  # the quantile is over all pairs: 501 counters summed across ranks (host logic, not a data-path collective)
            t = torch.from_numpy(hist.astype(np.int64)).cuda()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ghist = t.cpu().numpy().astype(np.uint64)
        thr, mc = C.c_double(0), C.c_int(0)
        check(L.dyna_quantile_type7_counts(ptr(ghist, C.c_uint64), n_hash, float(thresh_p), C.byref(thr), C.byref(mc)))
        cap = int(hist[max(mc.value, 1):].sum())
        if pin_ei is None or cap > edge_cap[0]:
            edge_cap[0] = cap
            pin_ei = torch.empty(max(cap, 1), dtype=torch.int32).pin_memory()
            pin_ej = torch.empty(max(cap, 1), dtype=torch.int32).pin_memory()
            pin_ec = torch.empty(max(cap, 1), dtype=torch.int16).pin_memory()
        ne = C.c_int64(0)
        check(L.dyna_mh_plan_threshold_edges(mplan, mc.value, cap, C.cast(pin_ei.data_ptr(), C.POINTER(C.c_int32)),
                                             C.cast(pin_ej.data_ptr(), C.POINTER(C.c_int32)),
                                             C.cast(pin_ec.data_ptr(), C.POINTER(C.c_uint16)), C.byref(ne), st))
        return thr.value, ne.value

    mh_sparse_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        sp_thr, sp_edges = mh_sparse_step()
    barrier()
    mh_sparse_s = max_over_ranks((time.perf_counter() - t0) / 3)
    sp_edges_total = sum_over_ranks(sp_edges)
    launches += 3 * (3 + 1 + 3)
    L.dyna_mh_plan_destroy(mplan)
    mh_alg_bytes = MH_BYTES_PER_PAIR * mh_total_pairs + 4.0 * mn * n_hash * world  # every rank reads all signatures once

    # ================================================================== CPU baseline (rank 0, N=1)
    cpu = None
    cpu_mh = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        gc, dt, cells, kind = cpu_nw_sample(seqs, 128)
        cpu = {"value": gc, "unit": "GCUPS", "cores": 1, "kind": kind,
               "sample": "reference similarityNW on the first 128 sequences of config 5 (8256 pairs, %.3g cells, %.1f s); "
                         "the reference NW is single-threaded" % (cells, dt)}
        cores = os.cpu_count() or 1
        rate, dt, pairs, kind = cpu_mh_sample(peps, 12000, cores)
        cpu_mh = {"value": rate, "unit": "pairs/s", "cores": cores, "kind": kind,
                  "sample": "reference similarityMH on the first 12000 peptides of config 4 (%d pairs, %.1f s), OpenMP on all host cores" % (pairs, dt)}
        rate1, dt1, pairs1, _ = cpu_mh_sample(peps, 4000, 1)
        cpu_mh["one_core"] = {"value": rate1, "unit": "pairs/s", "cores": 1,
                              "sample": "first 4000 peptides (%d pairs, %.1f s), OMP_NUM_THREADS=1" % (pairs1, dt1)}

    # ================================================================== BASELINE configs 1-3 through the drop-in API (rank 0)
    other = None
    if rank == 0:
        other = small_configs(dev, with_cpu=(world == 1 and not args.skip_cpu))

    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    tinfo = {}
    if os.path.exists(tpath):
        with open(tpath) as f:
            tinfo = json.load(f)
        # profiles/traffic.json holds two captures: the dominant launch of the FULL config-5 workload on one GPU (dram bytes
        # only) -- reported as `traffic` when this run is that workload -- and `ncu --set full` captures of a reduced
        # workload (tools/prof_target.py), reported next to it
        full = tinfo.get("nw_config5_dominant_launch")
        if full and world == 1 and n == 20000:
            traffic = full["dram_bytes"]

    if rank == 0:
        achieved = NW_OPS_PER_CELL * total_cells / (nw_ms_per_step * 1e-3) / world  # per GPU
        line = {
            "metric": "nw_allpairs_gcups", "value": nw_gcups, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": nw_ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "s16x2", "data": "synthetic",
            "dtype_note": "the reference's int32 DP evaluated exactly in packed 16-bit lanes (host range check per work unit; "
                          "units that could leave int16 run the int32 kernels)",
            "config": {"workload": "similarityNW BLOSUM62 gapOpen=10 gapExt=4 on synthetic %d proteins of ~330 aa (BASELINE config 5), "
                                   "all %d pairs i<=j = %.4g DP cells per step; row blocks balanced by cells over %d rank(s)"
                                   % (n, int(total_pairs), total_cells, world),
                       "l2": "256 MB flush write between timed steps; NW inputs (6.6 MB) are L2-resident by nature, outputs 8 B/pair",
                       "timing": "per-step CUDA events on the launching stream, max over ranks"},
            "clocks": clocks,
            "e2e": {"value": nw_e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": nw_h2d, "d2h_bytes_per_step": nw_d2h,
                    "api": "dyna_nw_pair_stats (validate + encode + plan + H2D + kernels + D2H), pinned host buffers"},
            "gpu_launches": launches,
            "roofline": {"bound": "int32_issue", "achieved": achieved / 1e9, "peak": int_peak / 1e9, "unit": "Gop/s",
                         "frac": achieved / int_peak, "traffic": traffic,
                         "traffic_note": ("dram__bytes_read.sum + dram__bytes_write.sum of the dominant launch (nw_warp2_kernel<11>, "
                                          "%.3g pairs of this workload), profiles/r01d_nw_config5_traffic.csv; algorithmic %.3g B"
                                          % (tinfo["nw_config5_dominant_launch"]["pairs_upper_bound"],
                                             tinfo["nw_config5_dominant_launch"]["algorithmic_bytes"])) if traffic else None,
                         "traffic_reduced_capture": {"dram_bytes": tinfo.get("nw_warp_kernel_dram_bytes_per_launch"),
                                                     "kernel": tinfo.get("nw_warp_kernel_dram_bytes_per_launch_kernel"),
                                                     "grid": tinfo.get("nw_warp_kernel_dram_bytes_per_launch_grid"),
                                                     "note": "ncu --set full, NW n=700 (0.27e11 cells): inputs stay in L2, DRAM traffic is negligible by construction"},
                         "note": "dominant kernel nw_warp2_kernel (16-bit DPX, two pairs per warp): neither HBM- nor tensor-bound; 11 algorithmic integer ops per DP cell "
                                 "(SURVEY.md 8(d)) against the INT32 issue peak measured live by dyna_probe_int_issue (IADD3 chains)"},
            "cpu_baseline": cpu,
            "minhash": {
                "metric": "minhash_pairs_per_sec", "value": mh_pairs_s, "unit": "pairs/s", "ms_per_step": mh_ms,
                "config": {"workload": "similarityMH k=4 n_hash=500 on synthetic %d peptides of 16 aa (BASELINE config 4), %d pairs per step; "
                                       "signatures rebuilt every step; u16 match counts for the strict upper triangle stay in HBM"
                                       % (mn, int(mh_total_pairs)),
                           "l2": "inputs (2 x %.0f MB signatures) and the %.1f GB output exceed the 126 MB L2" % (4.0 * mn * mh_hrows(n_hash) / 1e6, 2.0 * mh_total_pairs / 1e9)},
                "e2e": {"value": mh_total_pairs / mh_e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": int(mres.nbytes + moff.nbytes + seeds.nbytes),
                        "d2h_bytes_per_step": int(2 * mh_my_pairs), "api": "dyna_mh_plan_upload_sequences + run_signatures + run_match_fetch (chunked match, D2H overlapped)"},
                "exchange": ("relabelling sharded by code rows; one NCCL all-gather of the %.0f MB code table + max-reduce of the overflow gate per step"
                             % (msig.table.numel() * 4 / 1e6)) if msig.sharded else "none (every rank relabels all rows)",
                "e2e_sparse": {"value": mh_total_pairs / mh_sparse_s, "unit": "pairs/s", "thresh_p": 0.8, "threshold": sp_thr,
                               "edges": int(sp_edges_total), "d2h_bytes_per_step": int(10 * sp_edges + 8 * (n_hash + 1)),
                               "api": "upload_sequences + run_signatures + run_match + count_histogram + dyna_quantile_type7_counts + "
                                      "threshold_edges: clusterbreak's threshold step (R/clusterbreak.R:219-221) as an edge list, host buffers"},
                "roofline": {"bound": "hbm", "achieved": mh_alg_bytes / world / (mh_match_ms * 1e-3) / 1e9,
                             "peak": peaks_hbm(), "unit": "GB/s",
                             "frac": mh_alg_bytes / world / (mh_match_ms * 1e-3) / 1e9 / peaks_hbm(),
                             "traffic": (tinfo.get("mh_config4_match_launch", {}).get("dram_bytes")
                                         if world == 1 and mn == 100000 else None),
                             "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of the match launch at this workload (profiles/r01d_mh_config4_traffic.csv)",
                             "traffic_reduced_capture": {"dram_bytes": tinfo.get("mh_match_kernel_dram_bytes_per_launch"),
                                                         "algorithmic_bytes": 2.0 * 536854528 + 4.0 * 32768 * 500,
                                                         "note": "ncu --set full, MinHash n=32768 (536,854,528 pairs): measured DRAM bytes vs algorithmic"},
                             "note": "BASELINE names the HBM roofline (2 B/pair + signatures, peak = measured hbm_gbs); the kernel is "
                                     "integer-issue-bound by construction (n_hash compares per pair), see int32_issue_frac"},
                "match_kernel_ms": mh_match_ms,
                "int32_issue_frac": n_hash * mh_total_pairs / world / (mh_match_ms * 1e-3) / int_peak,
                "int32_issue_note": "n_hash equality compares per pair (1 lane-op each, algorithmic) per second / INT32 lane-op issue peak",
                "cpu_baseline": cpu_mh,
            },
            "int32_issue_peak_lane_ops_per_s": int_peak,
            "other_configs": other,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def small_configs(dev, with_cpu=False):
    """BASELINE configs 1, 2 and the sim_fn of config 3 on the reference's own data (fixtures under tests/golden),
    end to end through the drop-in entry points: host strings in, host n x n double matrix out."""
    import gzip

    import dynaalign_b200 as da
    gdir = os.path.join(ROOT, "tests", "golden")
    try:
        with open(os.path.join(gdir, "evp_probe_sequences.txt")) as f:
            evp = [ln.strip() for ln in f if ln.strip()]
        with gzip.open(os.path.join(gdir, "h3n2sample_first1000.json.gz"), "rt") as f:
            d = json.load(f)
        h3 = [d["unique"][i] for i in d["index"]]
    except OSError:
        return None

    def best_of(fn, reps=3):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return min(ts)

    out = {}
    t = best_of(lambda: da.similarityMH(evp, 2, 50, seed=42))
    out["config1_similarityMH_evp_k2_h50"] = {"n": len(evp), "seconds": t, "pairs_per_s": len(evp) * (len(evp) - 1) / 2 / t}
    lens = np.array([len(s) for s in h3], dtype=np.int64)
    cells = int((lens * np.cumsum(lens[::-1])[::-1]).sum())
    t = best_of(lambda: da.similarityNW(h3))
    out["config2_similarityNW_h3n2_1000"] = {"n": len(h3), "cells": cells, "seconds": t, "gcups": cells / t / 1e9}
    if with_cpu:  # the reference on a bounded sample of the same input (first 32 sequences), 1 core
        gc, dt, ccells, kind = cpu_nw_sample(h3, 32)
        out["config2_similarityNW_h3n2_1000"]["cpu_baseline"] = {
            "value": gc, "unit": "GCUPS", "cores": 1, "kind": kind,
            "sample": "reference similarityNW on h3n2sample[1:32] (%d cells, %.1f s)" % (ccells, dt)}
    t = best_of(lambda: da.similarityMH(h3, 4, 500, seed=42))
    out["config3_simfn_similarityMH_h3n2_1000_k4_h500"] = {"n": len(h3), "seconds": t,
                                                            "pairs_per_s": len(h3) * (len(h3) - 1) / 2 / t}
    # config 3 as a whole: the clusterbreak recursion on device-resident plans.  Louvain is igraph's (absent here), so the
    # clustering step is the deterministic connected-components stand-in and its share is reported separately.
    try:
        spent = [0.0]

        def timed_components(nv, gi, gj, gw):
            t0 = time.perf_counter()
            r = da.connected_components(nv, gi, gj, gw)
            spent[0] += time.perf_counter() - t0
            return r

        t0 = time.perf_counter()
        res = da.clusterbreak(h3, timed_components, thresh_p=0.8, size_max=800, size_min=3, max_itr=50, k=4, n_hash=500, seed=42,
                              verbose=False)
        total = time.perf_counter() - t0
        out["config3_clusterbreak_h3n2_1000"] = {
            "seconds_total": total, "seconds_cluster_fn_host": spent[0], "seconds_similarity_threshold_edges": total - spent[0],
            "recursion_nodes": res["calls"], "clustered": int(len(res["clustered_seq"])), "filtered": len(res["filtered_seq"]),
            "note": "size_max=800 thresh_p=0.8 sim_fn=similarityMH(k=4, n_hash=500); cluster_fn = connected components (igraph Louvain is "
                    "third-party and not installed); signatures hashed once, sub-clusters gather them on the device"}
    except Exception as e:
        out["config3_clusterbreak_h3n2_1000"] = {"error": str(e)[:200]}
    # BASELINE.json's stated target for NW: all pairs of the 100,000 config-4 peptides (5.00005e9 pairs, 1.28e12 cells),
    # device-resident plan, result left in HBM (40 GB as u32 matches + u32 length per pair)
    try:
        out["target_nw_100k_peptides_device"] = nw_peptides_target(dev)
    except Exception as e:  # never let a side measurement take the headline line down
        out["target_nw_100k_peptides_device"] = {"error": str(e)[:200]}
    out["note"] = ("wall clock of the drop-in call (flatten + validate + H2D + kernels + expansion to the column-major double matrix "
                   "+ D2H), best of 3; inputs are the reference's evp_peparray / h3n2sample extracts")
    return out


def nw_peptides_target(dev):
    import torch

    from dynaalign_b200 import synth
    from dynaalign_b200._lib import flatten, last_error, lib, ptr
    free, _ = torch.cuda.mem_get_info(dev)
    if free < 60e9:
        return {"skipped": "needs 40 GB of free device memory, found %.0f GB" % (free / 1e9)}
    seqs = synth.peptides_uniform(100_000)
    res, off = flatten(seqs)
    L = lib()
    plan = L.dyna_nw_plan_create(ptr(res, C.c_uint8), ptr(off, C.c_int64), len(seqs), b"BLOSUM62", 10, 4, 0, len(seqs), int(dev))
    if not plan:
        raise RuntimeError(last_error())
    try:
        cells, pairs = L.dyna_nw_plan_cells(plan), L.dyna_nw_plan_pairs(plan)
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            st = torch.cuda.current_stream(dev)
            e0.record(st)
            if L.dyna_nw_plan_run(plan, C.c_void_p(st.cuda_stream)) != 0:
                raise RuntimeError(last_error())
            e1.record(st)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) / 1e3)
    finally:
        L.dyna_nw_plan_destroy(plan)
    t = min(ts[1:])
    return {"n": len(seqs), "pairs": pairs, "cells": cells, "seconds": t, "gcups": cells / t / 1e9, "kernel": "nw_thread2_kernel"}


def mh_hrows(n_hash):
    return ((n_hash + 15) // 16) * 16


def peaks_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f).get("hbm_gbs", 6650.0))
    return 6650.0  # fallback stated in B200_PROFILING.md


if __name__ == "__main__":
    main()
