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
  * `e2e`: the R-facing call itself, dyna_similarityNW(sequences, "BLOSUM62", 10, 4, n_gpus = N), issued ONCE from
    rank 0, which drives all N GPUs of the box in-process (one host thread per device; the other ranks wait on a CPU
    barrier): pinned host buffers in, the whole n x n double matrix out, H2D + kernels + device-side gather of the
    slabs over NVLink + expansion + D2H inside the timed region.  `e2e_rowrange` is the same work through the
    multi-process row-range API (every rank calls dyna_nw_pair_stats for its own slab).
  * `parity_check`: position-weighted 64-bit checksums of every rank's (matches, length) slab and of the MinHash
    counts (plus the count histogram), summed over ranks and compared with tests/golden/bench_checksums.json
    (written by `bench.py --write-golden` on one GPU).  A partition that loses or repeats a row changes the sum.
    A mismatch makes the process exit non-zero.
  * `roofline`: for the dominant kernel (nw_rows2_kernel): algorithmic integer ops (11 per DP cell)
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
GOLDEN_CHECKSUMS = os.path.join(ROOT, "tests", "golden", "bench_checksums.json")
MASK64 = (1 << 64) - 1


def load_golden():
    if os.path.exists(GOLDEN_CHECKSUMS):
        with open(GOLDEN_CHECKSUMS) as f:
            return json.load(f)
    return {}


class Ctx:
    """Everything the phases share: ranks, streams, timing helpers."""

    def __init__(self, args, torch, dist, rank, world, local_rank):
        from dynaalign_b200 import _lib
        self.args, self.torch, self.dist = args, torch, dist
        self.rank, self.world, self.dev = rank, world, local_rank
        self._lib = _lib
        self.L = _lib.lib()
        self.stream = torch.cuda.current_stream()
        self.st = C.c_void_p(self.stream.cuda_stream)
        self.cpu_group = dist.new_group(backend="gloo") if world > 1 else None
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
        self.launches = 0

    def check(self, rc):
        self._lib.check(rc)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def cpu_barrier(self):
        """Ranks that only wait must not keep a NCCL kernel spinning on their GPU: rendezvous over gloo instead."""
        if self.world > 1:
            self.dist.barrier(group=self.cpu_group)

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather(self, obj):
        """list of every rank's `obj` (host objects, over the gloo group)."""
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.cpu_group)
        return out

    def timed_steps(self, fn, warmup, steps):
        """K steps, each bracketed by events on the launching stream; L2 flushed between steps (outside the events)."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for k in range(steps):
            self.flush_buf.fill_(k)
            e0[k].record(self.stream)
            fn()
            e1[k].record(self.stream)
        self.barrier()
        self.last_step_ms = [a.elapsed_time(b) for a, b in zip(e0, e1)]
        ms = sum(self.last_step_ms)
        return self.max_over_ranks(ms)

    def wall_steps(self, fn, steps):
        """Wall clock of `steps` calls between two barriers (host API calls that synchronise themselves), max over ranks."""
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.barrier()
        return self.max_over_ranks((time.perf_counter() - t0) / steps)

    def release_memory(self):
        self.torch.cuda.synchronize()
        self.torch.cuda.empty_cache()
        self.check(self.L.dyna_release_cached_memory(self.dev))


def pinned(torch, arr):
    return torch.from_numpy(arr).pin_memory()


def u8p(t):
    return C.cast(t.data_ptr(), C.POINTER(C.c_uint8))


# ================================================================== NW, config 5 (headline)
def phase_nw(ctx, sampler):
    from dynaalign_b200._lib import flatten, ptr
    L, args, torch = ctx.L, ctx.args, ctx.torch
    seqs = nw_workload(args.nw_n)
    n = len(seqs)
    res, off = flatten(seqs)
    lens = np.diff(off)
    bounds = np.zeros(ctx.world + 1, dtype=np.int64)
    ctx.check(L.dyna_partition_rows(n, ptr(lens.astype(np.int64), C.c_int64), 1, ctx.world, ptr(bounds, C.c_int64)))
    rb, re_ = int(bounds[ctx.rank]), int(bounds[ctx.rank + 1])
    plan = L.dyna_nw_plan_create(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, b"BLOSUM62", 10, 4, rb, re_, ctx.dev)
    if not plan:
        raise SystemExit("bench.py: " + ctx._lib.last_error())
    my_cells, my_pairs = L.dyna_nw_plan_cells(plan), L.dyna_nw_plan_pairs(plan)
    total_cells, total_pairs = ctx.sum_over_ranks(my_cells), ctx.sum_over_ranks(my_pairs)

    sampler.start()
    ms = ctx.timed_steps(lambda: ctx.check(L.dyna_nw_plan_run(plan, ctx.st)), args.warmup, args.steps)
    clocks = sampler.stop()
    per_step = L.dyna_nw_plan_launches(plan)
    ctx.launches += per_step * (args.steps + args.warmup)
    ms_per_step = ms / args.steps
    # parity: checksums of this rank's slab (the plan still holds the last step's result)
    h = (C.c_uint64 * 2)()
    ctx.check(L.dyna_nw_plan_checksum(plan, h, ctx.st))
    ctx.launches += 2
    sums = ctx.gather((int(h[0]), int(h[1])))
    checksum = [sum(x[0] for x in sums) & MASK64, sum(x[1] for x in sums) & MASK64]
    L.dyna_nw_plan_destroy(plan)

    # row-range e2e: every rank fetches its own slab through the host API (pinned buffers, H2D + D2H timed)
    pin_res, pin_off = pinned(torch, res), pinned(torch, off)
    out_m = torch.empty(max(my_pairs, 1), dtype=torch.int32).pin_memory()
    out_l = torch.empty(max(my_pairs, 1), dtype=torch.int32).pin_memory()
    ctx.check(L.dyna_set_device(ctx.dev))

    def rowrange_step():
        ctx.check(L.dyna_nw_pair_stats(u8p(pin_res), C.cast(pin_off.data_ptr(), C.POINTER(C.c_int64)), n, b"BLOSUM62", 10, 4,
                                       rb, re_, C.cast(out_m.data_ptr(), C.POINTER(C.c_uint32)),
                                       C.cast(out_l.data_ptr(), C.POINTER(C.c_uint32))))

    rowrange_step()  # warm-up (allocator, first touch)
    e2e_steps = max(1, min(args.steps, 3))
    rr_s = ctx.wall_steps(rowrange_step, e2e_steps)
    ctx.launches += per_step * (e2e_steps + 1)
    if my_pairs > 0:  # self-alignment of the slab's first row is an identity
        assert int(out_m[0]) == int(lens[rb]) and int(out_l[0]) == int(lens[rb]), "NW row-range e2e sanity check failed"
    first_row = (out_m[:n - rb].numpy().copy(), out_l[:n - rb].numpy().copy()) if my_pairs > 0 else None
    del out_m, out_l
    ctx.release_memory()
    return {"seqs": seqs, "n": n, "res": res, "off": off, "lens": lens, "rb": rb, "pin_res": pin_res, "pin_off": pin_off,
            "total_cells": total_cells, "total_pairs": total_pairs, "my_pairs": my_pairs, "ms_per_step": ms_per_step,
            "gcups": total_cells / (ms_per_step * 1e-3) / 1e9, "clocks": clocks, "launch_per_step": per_step,
            "checksum": checksum, "rowrange_gcups": total_cells / rr_s / 1e9, "rowrange_s": rr_s,
            "rowrange_d2h": int(8 * my_pairs), "h2d": int(res.nbytes + off.nbytes), "first_row": first_row}


# ================================================================== the R-facing call, all N GPUs driven from rank 0
def inproc_similarity_nw(ctx, seqs_flat, n, n_gpus, steps, check_row=None, median=False):
    """dyna_similarityNW(..., n_gpus) from rank 0 with pinned buffers; (seconds per call, sanity) or None elsewhere."""
    torch, L = ctx.torch, ctx.L
    res, off = seqs_flat
    ctx.release_memory()
    ctx.cpu_barrier()  # every rank has handed its cached device memory back
    out = None
    secs = None
    if ctx.rank == 0:
        pin_res, pin_off = pinned(torch, res), pinned(torch, off)
        out = torch.empty(n * n, dtype=torch.float64).pin_memory()

        def call():
            ctx.check(L.dyna_similarityNW(u8p(pin_res), C.cast(pin_off.data_ptr(), C.POINTER(C.c_int64)), n, b"BLOSUM62", 10, 4,
                                          C.cast(out.data_ptr(), C.POINTER(C.c_double)), n_gpus))

        call()  # warm-up: contexts, peer mappings, allocator pools of all devices
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            call()
            ts.append(time.perf_counter() - t0)
        secs = float(np.median(ts)) if median else float(np.mean(ts))
        m = out.numpy().reshape(n, n)  # column-major n x n: m[c, r] = matrix(r, c); symmetric by construction
        ok = bool(m[0, 0] == 1.0 and m[n - 1, n - 1] == 1.0 and m[3, n - 2] == m[n - 2, 3])
        if check_row is not None:  # row 0 of the matrix against the (matches, length) slab fetched earlier
            mt, ln = check_row
            ok = ok and bool((m[:, 0] == mt.astype(np.float64) / ln.astype(np.float64)).all() and (m[0, :] == m[:, 0]).all())
        if not ok:
            raise SystemExit("bench.py: in-process similarityNW result failed its sanity check")
        del out, pin_res, pin_off
        for d in range(n_gpus):
            ctx.check(L.dyna_release_cached_memory(d))
    ctx.cpu_barrier()
    return secs


# ================================================================== MinHash, config 4
def phase_mh(ctx):
    from dynaalign_b200._lib import flatten, ptr
    from dynaalign_b200.multirank import ShardedSignatures
    L, args, torch, dist = ctx.L, ctx.args, ctx.torch, ctx.dist
    peps = mh_workload(args.mh_n)
    mn, n_hash, k = len(peps), 500, 4
    mres, moff = flatten(peps)
    seeds = np.zeros(n_hash, dtype=np.uint32)
    ctx.check(L.dyna_hashfamily_seeds(42, n_hash, ptr(seeds, C.c_uint32)))
    mb = np.zeros(ctx.world + 1, dtype=np.int64)
    ctx.check(L.dyna_partition_rows(mn, None, 0, ctx.world, ptr(mb, C.c_int64)))
    mrb, mre = int(mb[ctx.rank]), int(mb[ctx.rank + 1])
    mplan = L.dyna_mh_plan_create(mn, n_hash, mrb, mre, ctx.dev)
    if not mplan:
        raise SystemExit("bench.py: " + ctx._lib.last_error())
    my_pairs = L.dyna_mh_plan_pairs(mplan)
    total_pairs = ctx.sum_over_ranks(my_pairs)
    ctx.check(L.dyna_mh_plan_upload_sequences(mplan, ptr(mres, C.c_uint8), ptr(moff, C.c_int64), k, ptr(seeds, C.c_uint32), ctx.st))
    # N > 1: the relabelling of the signature rows is sharded across ranks and completed by one NCCL all-gather of the
    # code table (+ a max-reduce of the overflow gate); N = 1 (or DYNA_MH_SHARD_RELABEL=0): plain run_signatures
    msig = ShardedSignatures(mplan, ctx.world if os.environ.get("DYNA_MH_SHARD_RELABEL", "1") != "0" else 1, ctx.rank, dist, torch,
                             torch.device("cuda", ctx.dev))

    def mh_step():
        msig.run(ctx.st)
        ctx.check(L.dyna_mh_plan_run_match(mplan, ctx.st))

    steps = max(args.steps, 3)
    mh_ms = ctx.timed_steps(mh_step, max(args.warmup, 3), steps) / steps
    ctx.launches += 3 * (steps + max(args.warmup, 3))
    msig.run(ctx.st)
    match_ms = ctx.timed_steps(lambda: ctx.check(L.dyna_mh_plan_run_match(mplan, ctx.st)), 1, 3) / 3
    ctx.launches += 4
    # parity: checksum + histogram of this rank's counts slab
    h = C.c_uint64(0)
    ctx.check(L.dyna_mh_plan_checksum(mplan, C.byref(h), ctx.st))
    hist = np.zeros(n_hash + 1, dtype=np.uint64)
    ctx.check(L.dyna_mh_plan_count_histogram(mplan, ptr(hist, C.c_uint64), ctx.st))
    ctx.launches += 2
    parts = ctx.gather((int(h.value), hist.tolist()))
    checksum = sum(x[0] for x in parts) & MASK64
    ghist = np.sum(np.array([x[1] for x in parts], dtype=np.uint64), axis=0)

    pin_mres, pin_moff = pinned(torch, mres), pinned(torch, moff)
    moffp = C.cast(pin_moff.data_ptr(), C.POINTER(C.c_int64))

    # dense e2e, narrow host form: 1 byte per pair + exact escapes for counts >= 255 (lossless; half the PCIe bytes of
    # the u16 triangle, which is what bounds this path)
    pin_c8 = torch.empty(max(my_pairs, 1), dtype=torch.uint8).pin_memory()
    esc_cap = 1 << 22
    esc_i = np.zeros(esc_cap, dtype=np.int64)
    esc_c = np.zeros(esc_cap, dtype=np.uint16)
    n_esc = C.c_int64(0)

    def dense8_step():
        ctx.check(L.dyna_mh_plan_upload_sequences(mplan, u8p(pin_mres), moffp, k, ptr(seeds, C.c_uint32), ctx.st))
        msig.run(ctx.st)
        ctx.check(L.dyna_mh_plan_run_match_fetch8(mplan, u8p(pin_c8), esc_cap, ptr(esc_i, C.c_int64), ptr(esc_c, C.c_uint16),
                                                  C.byref(n_esc), ctx.st))

    dense8_step()
    dense8_s = ctx.wall_steps(dense8_step, 3)
    ctx.launches += 4 * (3 + 16 * 2)
    # keep a bounded prefix of the narrow result (escapes restored) to compare with the u16 form below: lossless
    keep = int(min(my_pairs, 1 << 27))
    first = mrb * mn - mrb * (mrb + 1) // 2
    c8_prefix = pin_c8.numpy()[:keep].astype(np.uint16)
    ne = int(min(n_esc.value, esc_cap))
    sel = (esc_i[:ne] >= first) & (esc_i[:ne] < first + keep)
    c8_prefix[esc_i[:ne][sel] - first] = esc_c[:ne][sel]
    del pin_c8

    # dense e2e, u16 form (the plain triangle)
    pin_counts = torch.empty(max(my_pairs, 1), dtype=torch.int16).pin_memory()

    def dense16_step():
        ctx.check(L.dyna_mh_plan_upload_sequences(mplan, u8p(pin_mres), moffp, k, ptr(seeds, C.c_uint32), ctx.st))
        msig.run(ctx.st)
        ctx.check(L.dyna_mh_plan_run_match_fetch(mplan, C.cast(pin_counts.data_ptr(), C.POINTER(C.c_uint16)), ctx.st))

    dense16_step()
    dense16_s = ctx.wall_steps(dense16_step, 3)
    ctx.launches += 4 * (3 + 16)
    c8_ok = ctx.gather(bool((pin_counts.numpy()[:keep].view(np.uint16) == c8_prefix).all()))
    del pin_counts, c8_prefix

    # sparse e2e: what clusterbreak consumes at this size (the dense matrix would be 80 GB) -- the type-7 quantile threshold
    # and the edge list above it (R/clusterbreak.R:219-221).  Host buffers in, histogram + (i, j, count) edges out.
    shist = np.zeros(n_hash + 1, dtype=np.uint64)
    edge_cap = [0]
    pins = [None, None, None]

    j_inc, j_done = C.c_int64(0), C.c_int(0)

    def match_all_pairs():
        ctx.check(L.dyna_mh_plan_run_match(mplan, ctx.st))

    def match_join():  # csrc/mh_sparse.cu: join on equal signature values; all-pairs if it declines
        ctx.check(L.dyna_mh_plan_run_match_sparse(mplan, 0, C.byref(j_inc), C.byref(j_done), ctx.st))
        if not j_done.value:
            match_all_pairs()

    def sparse_step(thresh_p=0.8, matcher=match_all_pairs):
        ctx.check(L.dyna_mh_plan_upload_sequences(mplan, u8p(pin_mres), moffp, k, ptr(seeds, C.c_uint32), ctx.st))
        msig.run(ctx.st)
        matcher()
        ctx.check(L.dyna_mh_plan_count_histogram(mplan, ptr(shist, C.c_uint64), ctx.st))
        g = shist
        if ctx.world > 1:  # the quantile is over all pairs: 501 counters summed across ranks (host logic, not a data-path collective)
            t = torch.from_numpy(shist.astype(np.int64)).cuda()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            g = t.cpu().numpy().astype(np.uint64)
        thr, mc = C.c_double(0), C.c_int(0)
        ctx.check(L.dyna_quantile_type7_counts(ptr(g, C.c_uint64), n_hash, float(thresh_p), C.byref(thr), C.byref(mc)))
        cap = int(shist[max(mc.value, 1):].sum())
        if pins[0] is None or cap > edge_cap[0]:
            edge_cap[0] = cap
            pins[0] = torch.empty(max(cap, 1), dtype=torch.int32).pin_memory()
            pins[1] = torch.empty(max(cap, 1), dtype=torch.int32).pin_memory()
            pins[2] = torch.empty(max(cap, 1), dtype=torch.int16).pin_memory()
        ne = C.c_int64(0)
        ctx.check(L.dyna_mh_plan_threshold_edges(mplan, mc.value, cap, C.cast(pins[0].data_ptr(), C.POINTER(C.c_int32)),
                                                 C.cast(pins[1].data_ptr(), C.POINTER(C.c_int32)),
                                                 C.cast(pins[2].data_ptr(), C.POINTER(C.c_uint16)), C.byref(ne), ctx.st))
        return thr.value, ne.value

    sp = [sparse_step()]
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        sp.append(sparse_step())
    ctx.barrier()
    sparse_s = ctx.max_over_ranks((time.perf_counter() - t0) / 3)
    sp_thr, sp_edges = sp[-1]
    sp_edges_total = ctx.sum_over_ranks(sp_edges)
    ctx.launches += 4 * (3 + 1 + 3)
    table_mb = msig.table.numel() * 4 / 1e6 if msig.sharded else 0.0

    # the join (single GPU: it needs every hash row sorted on this device, which the sharded relabelling does not leave)
    join = None
    if ctx.world == 1 and not msig.sharded:
        def join_step():
            msig.run(ctx.st)
            match_join()

        # this step returns the incidence count to the host (a synchronising read inside the timed region), so host
        # hiccups land in it: the median step is reported, every step is listed
        ctx.timed_steps(join_step, 2, max(steps, 3))
        j_steps = list(ctx.last_step_ms)
        j_ms = float(np.median(j_steps))
        if j_done.value:
            jh = C.c_uint64(0)
            ctx.check(L.dyna_mh_plan_checksum(mplan, C.byref(jh), ctx.st))
            jhist = np.zeros(n_hash + 1, dtype=np.uint64)
            ctx.check(L.dyna_mh_plan_count_histogram(mplan, ptr(jhist, C.c_uint64), ctx.st))
            same = bool(int(jh.value) == checksum and (jhist == ghist).all())
            sj = [sparse_step(0.8, match_join)]
            t0 = time.perf_counter()
            for _ in range(3):
                sj.append(sparse_step(0.8, match_join))
            torch.cuda.synchronize()
            sj_s = (time.perf_counter() - t0) / 3
            join = {"ms": j_ms, "steps_ms": j_steps, "pairs_s": total_pairs / (j_ms * 1e-3), "incidences": int(j_inc.value), "same_as_all_pairs": same,
                    "sparse_s": sj_s, "sp_edges": int(sj[-1][1]), "sp_thr": sj[-1][0],
                    "same_edges": bool(sj[-1] == (sp_thr, sp_edges))}
            ctx.launches += (max(steps, 3) + 2) * 12 + 4 * 14
        else:
            join = {"declined": True, "incidences": int(j_inc.value)}
    L.dyna_mh_plan_destroy(mplan)
    ctx.release_memory()
    alg_bytes = MH_BYTES_PER_PAIR * total_pairs + 4.0 * mn * n_hash * ctx.world  # every rank reads all signatures once
    return {"peps": peps, "n": mn, "n_hash": n_hash, "total_pairs": total_pairs, "my_pairs": my_pairs, "ms": mh_ms,
            "match_ms": match_ms, "pairs_s": total_pairs / (mh_ms * 1e-3), "checksum": checksum, "hist": ghist,
            "dense8_s": dense8_s, "dense16_s": dense16_s, "n_esc": int(ctx.sum_over_ranks(n_esc.value)), "c8_ok": all(c8_ok),
            "sparse_s": sparse_s, "sp_thr": sp_thr, "sp_edges": sp_edges, "sp_edges_total": sp_edges_total,
            "h2d": int(mres.nbytes + moff.nbytes + seeds.nbytes), "sharded": msig.sharded, "table_mb": table_mb,
            "alg_bytes": alg_bytes, "join": join}


# ================================================================== NW on the 100,000 peptides (north_star target), all ranks
def safe_release(ctx):
    try:
        ctx.release_memory()
    except Exception:  # a failed side measurement may have left the device in an error state: report, do not die here
        pass


def phase_mixed_nw(ctx, kind="mix"):
    """NW on inputs whose rows are not all alike.  "mix": a proteome-like length mix (log-normal, median 300 residues, a
    few sequences beyond 1024) -- what the planner makes of rows that do not sit next to a row of their own length.
    "long": 1,500 proteins of 700..1020 residues -- the multi-pass kernels.  Rank 0 only; device time and an oracle sample."""
    if ctx.rank != 0:
        return None
    from dynaalign_b200._lib import flatten, ptr
    from dynaalign_b200 import synth
    L = ctx.L
    rng = np.random.default_rng(7)
    if kind == "mix":
        n = 3000
        lens = np.clip(rng.lognormal(np.log(300.0), 0.5, size=n), 30, 1800).astype(int)
    else:
        n = 1500
        lens = np.clip(np.rint(rng.normal(900.0, 60.0, size=n)), 700, 1020).astype(int)
    seqs = [synth.RESIDUES20[rng.integers(0, 20, size=int(x))].tobytes() for x in lens]
    res, off = flatten(seqs)
    plan = L.dyna_nw_plan_create(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, b"BLOSUM62", 10, 4, 0, n, ctx.dev)
    if not plan:
        raise RuntimeError(ctx._lib.last_error())
    try:
        cells = L.dyna_nw_plan_cells(plan)
        st = ctx.st
        ctx.check(L.dyna_nw_plan_run(plan, st))
        ev = [ctx.torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(ctx.stream)
        for _ in range(3):
            ctx.check(L.dyna_nw_plan_run(plan, st))
        ev[1].record(ctx.stream)
        ctx.torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 3
        launches = L.dyna_nw_plan_launches(plan)
        ctx.launches += 4 * launches
        pairs = L.dyna_nw_plan_pairs(plan)
        mt, ln = np.zeros(pairs, dtype=np.uint32), np.zeros(pairs, dtype=np.uint32)
        ctx.check(L.dyna_nw_plan_fetch(plan, ptr(mt, C.c_uint32), ptr(ln, C.c_uint32), None))
        ctx.torch.cuda.synchronize()
    finally:
        L.dyna_nw_plan_destroy(plan)
    ok = None
    if not ctx.args.skip_cpu:
        from oracle import port
        ok = True
        order = np.argsort(lens)
        picks = [(int(order[-1 - k]), int(order[-40 - k])) for k in range(6)]            # long against long
        picks += [(int(a), int(b)) for a, b in rng.integers(0, n, size=(150 if kind == "mix" else 40, 2))]
        for a, b in picks:
            i, j = min(a, b), max(a, b)
            slot = i * n - i * (i - 1) // 2 + (j - i)
            wm, wl = port.nw_pair(seqs[i], seqs[j])[:2]
            ok = ok and int(mt[slot]) == int(wm) and int(ln[slot]) == int(wl)
    ctx.release_memory()
    return {"n": n, "length_min_median_max": [int(lens.min()), int(np.median(lens)), int(lens.max())], "cells": int(cells),
            "seconds": ms * 1e-3, "gcups": cells / ms / 1e6, "kernel_launches_per_step": int(launches), "oracle_sample_ok": ok,
            "note": ("log-normal lengths (median 300, sigma 0.5, clipped to 30..1800), uniform residues, input order random: "
                     "two-rows partner search in a 32-row window, kernel choice per unit, device time of dyna_nw_plan_run")
                    if kind == "mix" else
                    "lengths Normal(900, 60) clipped to 700..1020, uniform residues: rows of 769+ residues through the two-rows "
                    "multi-pass kernel, the rest through the cooperative one; device time of dyna_nw_plan_run"}


def phase_target_nw(ctx):
    """All pairs of the config-4 peptides through NW, row blocks over the ranks: device time, e2e in the 2-bytes-per-pair
    host form, checksums, and a sample of pairs against the oracle (rank 0)."""
    from dynaalign_b200._lib import flatten, ptr
    L, torch = ctx.L, ctx.torch
    seqs = mh_workload(ctx.args.mh_n)
    n = len(seqs)
    free, _ = torch.cuda.mem_get_info(ctx.dev)
    bounds = np.zeros(ctx.world + 1, dtype=np.int64)
    lens = np.full(n, 16, dtype=np.int64)
    res, off = flatten(seqs)
    ctx.check(L.dyna_partition_rows(n, ptr(lens, C.c_int64), 1, ctx.world, ptr(bounds, C.c_int64)))
    rb, re_ = int(bounds[ctx.rank]), int(bounds[ctx.rank + 1])
    my_pairs_est = (re_ - rb) * n  # upper bound
    need = 10.0 * my_pairs_est + 2e9
    fits = all(ctx.gather(bool(free > need)))
    if not fits:
        return {"skipped": "needs %.0f GB of free device memory per rank" % (need / 1e9)}
    plan = L.dyna_nw_plan_create(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, b"BLOSUM62", 10, 4, rb, re_, ctx.dev)
    if not plan:
        raise RuntimeError(ctx._lib.last_error())
    try:
        my_pairs, my_cells = L.dyna_nw_plan_pairs(plan), L.dyna_nw_plan_cells(plan)
        cells, pairs = ctx.sum_over_ranks(my_cells), ctx.sum_over_ranks(my_pairs)
        ms = ctx.timed_steps(lambda: ctx.check(L.dyna_nw_plan_run(plan, ctx.st)), 1, 3) / 3
        per_step = L.dyna_nw_plan_launches(plan)
        ctx.launches += 4 * per_step
        h = (C.c_uint64 * 2)()
        ctx.check(L.dyna_nw_plan_checksum(plan, h, ctx.st))
        sums = ctx.gather((int(h[0]), int(h[1])))
        checksum = [sum(x[0] for x in sums) & MASK64, sum(x[1] for x in sums) & MASK64]
    finally:
        L.dyna_nw_plan_destroy(plan)
    # e2e: host strings in (pinned), (matches, length) as one byte each out (pinned): validate + plan + H2D + kernel + pack + D2H
    pin_res, pin_off = pinned(torch, res), pinned(torch, off)
    out_m = torch.empty(max(my_pairs, 1), dtype=torch.uint8).pin_memory()
    out_l = torch.empty(max(my_pairs, 1), dtype=torch.uint8).pin_memory()
    ctx.check(L.dyna_set_device(ctx.dev))

    def step():
        ctx.check(L.dyna_nw_pair_stats8(u8p(pin_res), C.cast(pin_off.data_ptr(), C.POINTER(C.c_int64)), n, b"BLOSUM62", 10, 4, rb, re_,
                                        u8p(out_m), u8p(out_l)))

    step()
    e2e_s = ctx.wall_steps(step, 2)
    ctx.launches += 3 * (per_step + 1)
    # parity sample (rank 0): 400 pairs of its slab against the oracle port, pair by pair
    sample_ok = None
    if ctx.rank == 0 and not ctx.args.skip_cpu:
        from oracle import port
        rng = np.random.default_rng(5)
        m8, l8 = out_m.numpy(), out_l.numpy()
        sample_ok = True
        for _ in range(400):
            i = int(rng.integers(rb, min(re_, rb + 50)))
            j = int(rng.integers(i, n))
            slot = (i * n - i * (i - 1) // 2 + (j - i)) - (rb * n - rb * (rb - 1) // 2)
            wm, wl = port.nw_pair(seqs[i], seqs[j])[:2]
            sample_ok = sample_ok and int(m8[slot]) == int(wm) and int(l8[slot]) == int(wl)
    del out_m, out_l
    ctx.release_memory()

    # sparse e2e: what clusterbreak consumes at this size -- the exact type-7 quantile of the identities over ALL pairs
    # and the edge list above it (R/clusterbreak.R:219-221 with sim_fn = similarityNW).  Host strings in, (matches, length)
    # histogram + (i, j, matches, length) edges out.  thresh_p = 0.999: NW identities are almost never 0, so the
    # reference's default 0.8 would keep a billion edges here; the top 0.1 % is ~5 M.
    from dynaalign_b200.api import identities_at_least
    offp = C.cast(pin_off.data_ptr(), C.POINTER(C.c_int64))
    pins = [None] * 4
    edge_cap = [0]
    SP_P = 0.999

    def sparse_step():
        p = L.dyna_nw_plan_create(u8p(pin_res), offp, n, b"BLOSUM62", 10, 4, rb, re_, ctx.dev)
        if not p:
            raise RuntimeError(ctx._lib.last_error())
        try:
            ctx.check(L.dyna_nw_plan_run(p, ctx.st))
            ml = L.dyna_nw_plan_max_len(p)
            hist = np.zeros((ml + 1, 2 * ml + 1), dtype=np.uint64)
            ctx.check(L.dyna_nw_plan_stat_histogram(p, None, 0, ptr(hist, C.c_uint64), ctx.st))
            g = hist
            if ctx.world > 1:  # the quantile is over all pairs: a few hundred counters summed across ranks (host logic)
                t = torch.from_numpy(hist.astype(np.int64)).cuda()
                ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.SUM)
                g = np.ascontiguousarray(t.cpu().numpy().astype(np.uint64))
            thr = C.c_double(0)
            ctx.check(L.dyna_quantile_type7_identities(ptr(g, C.c_uint64), ml + 1, 2 * ml + 1, SP_P, C.byref(thr)))
            cap = identities_at_least(hist, thr.value)
            if pins[0] is None or cap > edge_cap[0]:
                edge_cap[0] = cap
                for q in range(4):
                    pins[q] = torch.empty(max(cap, 1), dtype=torch.int32).pin_memory()
            ne = C.c_int64(0)
            ctx.check(L.dyna_nw_plan_threshold_edges(p, None, 0, thr.value, cap, C.cast(pins[0].data_ptr(), C.POINTER(C.c_int32)),
                                                     C.cast(pins[1].data_ptr(), C.POINTER(C.c_int32)),
                                                     C.cast(pins[2].data_ptr(), C.POINTER(C.c_uint32)),
                                                     C.cast(pins[3].data_ptr(), C.POINTER(C.c_uint32)), C.byref(ne), ctx.st))
        finally:
            L.dyna_nw_plan_destroy(p)
        return thr.value, ne.value, cap, identities_at_least(g, thr.value)

    sparse_step()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        sp_thr, sp_ne, sp_cap, sp_all = sparse_step()
    ctx.barrier()
    sparse_s = ctx.max_over_ranks((time.perf_counter() - t0) / 2)
    ctx.launches += 3 * (per_step + 4)
    sp_total = ctx.sum_over_ranks(sp_ne)
    sp_ok = all(ctx.gather(bool(sp_ne == sp_cap))) and sp_total == sp_all  # edge list == what the histogram says it must be
    if ctx.rank == 0 and not ctx.args.skip_cpu and sp_ne > 0:  # and a sample of the edges against the oracle port
        from oracle import port
        rng = np.random.default_rng(6)
        ei, ej, em, el = (pins[q].numpy()[:sp_ne] for q in range(4))
        for q in rng.integers(0, sp_ne, size=200):
            wm, wl = port.nw_pair(seqs[int(ei[q])], seqs[int(ej[q])])[:2]
            sp_ok = sp_ok and int(em[q]) == int(wm) and int(el[q]) == int(wl) and wm / wl >= sp_thr and ei[q] < ej[q]
    pins[:] = [None] * 4
    ctx.release_memory()
    return {"n": n, "pairs": int(pairs), "cells": int(cells), "seconds": ms * 1e-3, "gcups": cells / (ms * 1e-3) / 1e9,
            "kernel": "nw_thread_rows2_kernel", "checksum": checksum,
            "e2e": {"seconds": e2e_s, "gcups": cells / e2e_s / 1e9, "d2h_bytes_per_step": int(2 * my_pairs),
                    "api": "dyna_nw_pair_stats8 per rank: validate + encode + plan + H2D + kernel + u8 pack + D2H of 2 bytes per pair "
                           "(matches, length <= 32), pinned host buffers"},
            "e2e_sparse": {"seconds": sparse_s, "gcups": cells / sparse_s / 1e9, "thresh_p": SP_P, "threshold": sp_thr,
                           "edges": int(sp_total), "d2h_bytes_per_step": int(16 * sp_ne), "edges_consistent": bool(sp_ok),
                           "api": "dyna_nw_plan_create + run + stat_histogram + dyna_quantile_type7_identities + threshold_edges per rank: "
                                  "clusterbreak's threshold step (R/clusterbreak.R:219-221) for sim_fn = similarityNW as an edge list, "
                                  "host buffers in and out"},
            "oracle_sample_ok": sample_ok}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--nw-n", type=int, default=20000, help="config 5 size (development: smaller)")
    ap.add_argument("--mh-n", type=int, default=100000, help="config 4 size (development: smaller)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--write-golden", action="store_true", help="N=1 only: store this run's checksums as the parity goldens")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist

    from dynaalign_b200 import _lib

    L = _lib.lib()
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
    ctx = Ctx(args, torch, dist, rank, world, local_rank)
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    phys = vis.split(",")[local_rank] if vis else str(local_rank)

    # ---------------- integer issue peak (roofline denominator), measured live on this device
    peak_ops, peak_ms = C.c_double(0), C.c_double(0)
    ctx.check(L.dyna_probe_int_issue(0, C.byref(peak_ops), C.byref(peak_ms), ctx.st))
    int_peak = peak_ops.value  # lane-ops/s

    nw = phase_nw(ctx, ClockSampler(phys))
    n = nw["n"]
    mh = phase_mh(ctx)
    # ---------------- the R-facing call on all N GPUs of the box, from rank 0 (headline e2e)
    e2e_steps = max(1, min(args.steps, 3))
    inproc_s = inproc_similarity_nw(ctx, (nw["res"], nw["off"]), n, world, e2e_steps,
                                    check_row=nw["first_row"] if rank == 0 and nw["rb"] == 0 else None)
    inproc_c2 = None
    h3 = load_h3n2()
    if h3 is not None:
        from dynaalign_b200._lib import flatten
        c2_flat = flatten(h3)
        c2_s = inproc_similarity_nw(ctx, c2_flat, len(h3), world, 5, median=True)  # a 40 ms call: median of 5
        if rank == 0:
            lens2 = np.diff(c2_flat[1])
            cells2 = int((lens2 * np.cumsum(lens2[::-1])[::-1]).sum())
            inproc_c2 = {"n": len(h3), "n_gpus": world, "cells": cells2, "seconds": c2_s, "seconds_is": "median of 5 calls",
                         "gcups": cells2 / c2_s / 1e9}

    # ---------------- side measurements, after the headline numbers are in (a failure here costs only its own entry)
    try:
        target = phase_target_nw(ctx)
    except Exception as e:  # never let a side measurement take the headline line down
        target = {"error": str(e)[:200]}
        safe_release(ctx)
    try:
        mixed = phase_mixed_nw(ctx)
    except Exception as e:
        mixed = {"error": str(e)[:200]}
        safe_release(ctx)
    try:
        longp = phase_mixed_nw(ctx, "long")
    except Exception as e:
        longp = {"error": str(e)[:200]}
        safe_release(ctx)

    # ================================================================== CPU baseline (rank 0, N=1)
    cpu = None
    cpu_mh = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        gc, dt, cells, kind = cpu_nw_sample(nw["seqs"], 128)
        cpu = {"value": gc, "unit": "GCUPS", "cores": 1, "kind": kind,
               "sample": "reference similarityNW on the first 128 sequences of config 5 (8256 pairs, %.3g cells, %.1f s); "
                         "the reference NW is single-threaded" % (cells, dt)}
        cores = os.cpu_count() or 1
        rate, dt, pairs, kind = cpu_mh_sample(mh["peps"], 12000, cores)
        cpu_mh = {"value": rate, "unit": "pairs/s", "cores": cores, "kind": kind,
                  "sample": "reference similarityMH on the first 12000 peptides of config 4 (%d pairs, %.1f s), OpenMP on all host cores" % (pairs, dt)}
        rate1, dt1, pairs1, _ = cpu_mh_sample(mh["peps"], 4000, 1)
        cpu_mh["one_core"] = {"value": rate1, "unit": "pairs/s", "cores": 1,
                              "sample": "first 4000 peptides (%d pairs, %.1f s), OMP_NUM_THREADS=1" % (pairs1, dt1)}

    # ================================================================== BASELINE configs 1-3 through the drop-in API (rank 0)
    other = None
    if rank == 0:
        try:
            other = small_configs(local_rank, with_cpu=(world == 1 and not args.skip_cpu))
        except Exception as e:  # e.g. the device left in an error state by a failed side measurement above
            other = {"error": str(e)[:200]}
        if other is not None:
            other["target_nw_100k_peptides"] = target
            other["nw_mixed_lengths_3000"] = mixed
            other["nw_long_proteins_1500"] = longp
            other["similarityNW_inproc_n%d" % world] = {
                "config5": {"n": n, "n_gpus": world, "seconds": inproc_s, "gcups": nw["total_cells"] / inproc_s / 1e9,
                            "d2h_bytes": 8 * n * n},
                "config2": inproc_c2,
                "api": "dyna_similarityNW(residues, offsets, n, 'BLOSUM62', 10, 4, out, n_gpus) called once from one process: one "
                       "host thread per device, row blocks balanced by DP cells, column blocks of the n x n matrix expanded on "
                       "every device from all devices' slabs (peer loads over NVLink), one contiguous D2H per device"}
    ctx.cpu_barrier()

    # ================================================================== parity against the committed goldens
    golden = load_golden()
    key_nw = "nw_config5_n%d" % n
    key_mh = "mh_config4_n%d_k4_h%d_seed42" % (mh["n"], mh["n_hash"])
    key_tg = "nw_target_n%d" % mh["n"]
    if args.write_golden and rank == 0:
        if world != 1:
            raise SystemExit("bench.py: --write-golden needs a single-GPU run")
        golden[key_nw] = {"matches": nw["checksum"][0], "length": nw["checksum"][1], "pairs": int(nw["total_pairs"])}
        golden[key_mh] = {"counts": mh["checksum"], "hist": [int(x) for x in mh["hist"]], "pairs": int(mh["total_pairs"])}
        if "checksum" in target:
            golden[key_tg] = {"matches": target["checksum"][0], "length": target["checksum"][1], "pairs": target["pairs"],
                              "oracle_sample_ok": target.get("oracle_sample_ok")}
        golden["_about"] = ("position-weighted 64-bit checksums (dynaalign_b200.api.checksum) of the packed result triangles, "
                            "written by `python bench.py --write-golden` on ONE B200; bench.py compares the sum over all "
                            "ranks' slabs with these at every N")
        with open(GOLDEN_CHECKSUMS, "w") as f:
            json.dump(golden, f, indent=1, sort_keys=True)

    def cmp(key, got):
        want = golden.get(key)
        if want is None:
            return None
        return all(want[k] == v for k, v in got.items())

    parity = {
        "nw": cmp(key_nw, {"matches": nw["checksum"][0], "length": nw["checksum"][1]}),
        "mh": cmp(key_mh, {"counts": mh["checksum"], "hist": [int(x) for x in mh["hist"]]}),
        "mh_narrow_form_lossless": mh["c8_ok"],
        "target_nw": cmp(key_tg, {"matches": target["checksum"][0], "length": target["checksum"][1]}) if "checksum" in target else None,
        "target_nw_oracle_sample": target.get("oracle_sample_ok"),
        "nw_mixed_lengths_oracle_sample": (mixed or {}).get("oracle_sample_ok"),
        "nw_long_proteins_oracle_sample": (longp or {}).get("oracle_sample_ok"),
        "nw_checksum": ["%016x" % c for c in nw["checksum"]], "mh_checksum": "%016x" % mh["checksum"],
        "golden": os.path.relpath(GOLDEN_CHECKSUMS, ROOT) if golden else None,
        "note": "sum over all ranks' slabs of value[k] * w(global pair index k) mod 2^64 (+ the MinHash count histogram) against "
                "the single-GPU goldens; null = no golden for this workload size",
    }
    parity_failed = any(v is False for v in parity.values())

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
    kmet = tinfo.get("nw_config5_dominant_kernel_metrics", {})

    if rank == 0:
        nw_ms_per_step = nw["ms_per_step"]
        total_cells, total_pairs = nw["total_cells"], nw["total_pairs"]
        achieved = NW_OPS_PER_CELL * total_cells / (nw_ms_per_step * 1e-3) / world  # per GPU
        mn, n_hash, mh_total_pairs, mh_match_ms = mh["n"], mh["n_hash"], mh["total_pairs"], mh["match_ms"]
        line = {
            "metric": "nw_allpairs_gcups", "value": nw["gcups"], "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": nw_ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "s16x2", "data": "synthetic",
            "dtype_note": "the reference's int32 DP evaluated exactly in packed 16-bit lanes (host range check per work unit; "
                          "units that could leave int16 run the int32 kernels)",
            "config": {"workload": "similarityNW BLOSUM62 gapOpen=10 gapExt=4 on synthetic %d proteins of ~330 aa (BASELINE config 5), "
                                   "all %d pairs i<=j = %.4g DP cells per step; row blocks balanced by cells over %d rank(s)"
                                   % (n, int(total_pairs), total_cells, world),
                       "l2": "256 MB flush write between timed steps; NW inputs (6.6 MB) are L2-resident by nature, outputs 8 B/pair",
                       "timing": "per-step CUDA events on the launching stream, max over ranks"},
            "clocks": nw["clocks"],
            "e2e": {"value": total_cells / inproc_s / 1e9 if inproc_s else None, "unit": "GCUPS",
                    "h2d_bytes_per_step": nw["h2d"] * world, "d2h_bytes_per_step": 8 * n * n,
                    "api": "dyna_similarityNW(..., n_gpus=%d), the R-facing call, once from rank 0 driving all %d GPU(s) in-process: "
                           "validate + encode + plans + H2D + kernels + slab gather over NVLink peer loads + expansion to the "
                           "column-major n x n double matrix + D2H; pinned host buffers" % (world, world)},
            "e2e_rowrange": {"value": nw["rowrange_gcups"], "unit": "GCUPS", "h2d_bytes_per_step": nw["h2d"],
                             "d2h_bytes_per_step": nw["rowrange_d2h"],
                             "api": "dyna_nw_pair_stats per rank (validate + encode + plan + H2D + kernels + D2H of the rank's "
                                    "(matches, length) slab), pinned host buffers, max over ranks"},
            "gpu_launches": ctx.launches,
            "parity_check": parity,
            "roofline": {"bound": "int32_issue", "achieved": achieved / 1e9, "peak": int_peak / 1e9, "unit": "Gop/s",
                         "frac": achieved / int_peak, "traffic": traffic,
                         "issue_active_pct": kmet.get("smsp__issue_active_pct"), "alu_pipe_pct": kmet.get("sm__pipe_alu_pct"),
                         "shared_wavefronts_pct": kmet.get("l1tex_shared_wavefronts_pct"),
                         "utilisation_note": kmet.get("note"),
                         "traffic_note": ("dram__bytes_read.sum + dram__bytes_write.sum of the dominant launch (%s, <= %.3g pairs of this "
                                          "workload), %s; algorithmic %.3g B"
                                          % (tinfo["nw_config5_dominant_launch"]["kernel"], tinfo["nw_config5_dominant_launch"]["pairs_upper_bound"],
                                             tinfo["nw_config5_dominant_launch"]["source"].split(":")[0],
                                             tinfo["nw_config5_dominant_launch"]["algorithmic_bytes"])) if traffic else None,
                         "traffic_reduced_capture": {"dram_bytes": tinfo.get("nw_warp_kernel_dram_bytes_per_launch"),
                                                     "kernel": tinfo.get("nw_warp_kernel_dram_bytes_per_launch_kernel"),
                                                     "grid": tinfo.get("nw_warp_kernel_dram_bytes_per_launch_grid"),
                                                     "note": "ncu --set full, NW n=700 (0.27e11 cells): inputs stay in L2, DRAM traffic is negligible by construction"},
                         "note": "dominant kernel nw_rows2_kernel (16-bit DPX, two row sequences per warp against one column sequence): neither HBM- nor tensor-bound; 11 algorithmic integer ops per DP cell "
                                 "(SURVEY.md 8(d)) against the INT32 issue peak measured live by dyna_probe_int_issue (IADD3 chains); the packed kernel "
                                 "updates two cells per instruction, so frac is not a ceiling -- issue_active_pct / alu_pipe_pct are the utilisation"},
            "cpu_baseline": cpu,
            "minhash_pairs_per_sec": mh["pairs_s"],
            "minhash_e2e_pairs_per_sec": mh_total_pairs / mh["dense8_s"],
            "minhash_join_pairs_per_sec": (mh["join"] or {}).get("pairs_s"),
            "minhash": {
                "metric": "minhash_pairs_per_sec", "value": mh["pairs_s"], "unit": "pairs/s", "ms_per_step": mh["ms"],
                "config": {"workload": "similarityMH k=4 n_hash=500 on synthetic %d peptides of 16 aa (BASELINE config 4), %d pairs per step; "
                                       "signatures rebuilt every step; u16 match counts for the strict upper triangle stay in HBM"
                                       % (mn, int(mh_total_pairs)),
                           "l2": "inputs (2 x %.0f MB signatures) and the %.1f GB output exceed the 126 MB L2" % (4.0 * mn * mh_hrows(n_hash) / 1e6, 2.0 * mh_total_pairs / 1e9)},
                "e2e": {"value": mh_total_pairs / mh["dense8_s"], "unit": "pairs/s", "h2d_bytes_per_step": mh["h2d"],
                        "d2h_bytes_per_step": int(mh["my_pairs"]), "escapes": mh["n_esc"],
                        "d2h_gb_per_s_per_rank": mh["my_pairs"] / mh["dense8_s"] / 1e9,
                        "api": "dyna_mh_plan_upload_sequences + run_signatures + run_match_fetch8: chunked match, each chunk narrowed to one byte "
                               "per pair (+ exact escape list for counts >= 255: lossless) and copied while the next chunk is matched"},
                "e2e_u16": {"value": mh_total_pairs / mh["dense16_s"], "unit": "pairs/s", "d2h_bytes_per_step": int(2 * mh["my_pairs"]),
                            "api": "... + run_match_fetch (the plain u16 triangle)"},
                "exchange": ("relabelling sharded by code rows; one NCCL all-gather of the %.0f MB code table + max-reduce of the overflow gate per step"
                             % mh["table_mb"]) if mh["sharded"] else "none (every rank relabels all rows)",
                "e2e_sparse": {"value": mh_total_pairs / mh["sparse_s"], "unit": "pairs/s", "thresh_p": 0.8, "threshold": mh["sp_thr"],
                               "edges": int(mh["sp_edges_total"]), "d2h_bytes_per_step": int(10 * mh["sp_edges"] + 8 * (n_hash + 1)),
                               "api": "upload_sequences + run_signatures + run_match + count_histogram + dyna_quantile_type7_counts + "
                                      "threshold_edges: clusterbreak's threshold step (R/clusterbreak.R:219-221) as an edge list, host buffers"},
                "join": (None if mh["join"] is None else mh["join"] if "declined" in mh["join"] else {
                    "value": mh["join"]["pairs_s"], "unit": "pairs/s", "ms_per_step": mh["join"]["ms"],
                    "ms_per_step_is": "median of the listed steps", "steps_ms": mh["join"]["steps_ms"],
                    "incidences": mh["join"]["incidences"], "same_checksum_and_histogram_as_all_pairs": mh["join"]["same_as_all_pairs"],
                    "e2e_sparse": {"value": mh_total_pairs / mh["join"]["sparse_s"], "unit": "pairs/s", "thresh_p": 0.8,
                                   "threshold": mh["join"]["sp_thr"], "edges": mh["join"]["sp_edges"],
                                   "same_edges_as_all_pairs": mh["join"]["same_edges"]},
                    "note": "same step as `value` (signatures + relabelling + match counts, device-timed) with the match counts "
                            "computed by dyna_mh_plan_run_match_sparse: the hash rows are already sorted by the relabelling, every "
                            "group of equal signature values emits its pairs, the pair list is radix-sorted (CUB) and run-length "
                            "encoded.  Exact and data dependent: O(n * n_hash + matches) instead of O(n^2 * n_hash); `value` above stays "
                            "the all-pairs kernel, whose rate does not depend on the data.  Single GPU (N > 1 shards the all-pairs kernel)"}),
                "roofline": {"bound": "hbm", "achieved": mh["alg_bytes"] / world / (mh_match_ms * 1e-3) / 1e9,
                             "peak": peaks_hbm(), "unit": "GB/s",
                             "frac": mh["alg_bytes"] / world / (mh_match_ms * 1e-3) / 1e9 / peaks_hbm(),
                             "traffic": (tinfo.get("mh_config4_match_launch", {}).get("dram_bytes")
                                         if world == 1 and mn == 100000 else None),
                             "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of the match launch at this workload (profiles/r01d_mh_config4_traffic.csv): "
                                             "14.2 GB against 10.1 GB algorithmic -- 3.4 GB of re-reads of the 0.2 GB code table across tile groups and 0.8 GB of "
                                             "partial-sector writes at row ends; not binding (2.5 % of HBM), recorded as waste",
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
    if parity_failed:
        raise SystemExit("bench.py: parity check against %s FAILED: %s" % (GOLDEN_CHECKSUMS, {k: v for k, v in parity.items() if v is False}))


def load_h3n2():
    import gzip
    try:
        with gzip.open(os.path.join(ROOT, "tests", "golden", "h3n2sample_first1000.json.gz"), "rt") as f:
            d = json.load(f)
        return [d["unique"][i] for i in d["index"]]
    except OSError:
        return None


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
    # the same caller loop with sim_fn = similarityNW (config 2's similarity): aligned once, every recursion node read off
    # the root's triangle on the device; the reference re-aligns each node (R/clusterbreak.R:217,250-254)
    try:
        spent = [0.0]
        t0 = time.perf_counter()
        res = da.clusterbreak(h3, timed_components, thresh_p=0.8, size_max=800, size_min=3, max_itr=50, verbose=False, sim="NW")
        total = time.perf_counter() - t0
        out["config2_clusterbreak_similarityNW_h3n2_1000"] = {
            "seconds_total": total, "seconds_cluster_fn_host": spent[0], "seconds_similarity_threshold_edges": total - spent[0],
            "recursion_nodes": res["calls"], "clustered": int(len(res["clustered_seq"])), "filtered": len(res["filtered_seq"]),
            "note": "size_max=800 thresh_p=0.8 sim_fn=similarityNW(BLOSUM62, 10, 4); one NW triangle for all recursion nodes "
                    "((matches, length) histogram -> exact type-7 quantile -> edge list per node on the device)"}
    except Exception as e:
        out["config2_clusterbreak_similarityNW_h3n2_1000"] = {"error": str(e)[:200]}
    out["note"] =("wall clock of the drop-in call (flatten + validate + H2D + kernels + expansion to the column-major double matrix "
                   "+ D2H), best of 3; inputs are the reference's evp_peparray / h3n2sample extracts")
    return out


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
