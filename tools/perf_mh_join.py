"""Development helper (GPU box): BASELINE config 4 (and its clustered variant) through the join and through the all-pairs
kernel: wall time per phase, incidences, checksums.  python tools/perf_mh_join.py [n] [uniform|clustered]"""
import ctypes as C, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynaalign_b200 import synth
from dynaalign_b200._lib import check, flatten, lib, ptr
L = lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
kind = sys.argv[2] if len(sys.argv) > 2 else "uniform"
seqs = synth.peptides_uniform(n) if kind == "uniform" else synth.peptides_clustered(n)
res, off = flatten(seqs)
n_hash, k = 500, 4
seeds = np.zeros(n_hash, dtype=np.uint32)
check(L.dyna_hashfamily_seeds(42, n_hash, ptr(seeds, C.c_uint32)))
def wall(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3
p = L.dyna_mh_plan_create(n, n_hash, 0, n, 0)
check(L.dyna_mh_plan_upload_sequences(p, ptr(res, C.c_uint8), ptr(off, C.c_int64), k, ptr(seeds, C.c_uint32), None))
t_sig = wall(lambda: check(L.dyna_mh_plan_run_signatures(p, None)))
inc, done = C.c_int64(0), C.c_int(0)
cap = int(os.environ.get("JOIN_CAP", "0"))
t_join = wall(lambda: check(L.dyna_mh_plan_run_match_sparse(p, cap, C.byref(inc), C.byref(done), None)))
pairs = L.dyna_mh_plan_pairs(p)
print("n=%d %s: signatures %.2f ms, join %.2f ms (done=%d, %d incidences) -> %.3g pairs/s" % (n, kind, t_sig, t_join, done.value, inc.value, pairs / ((t_sig + t_join) * 1e-3)))
h = C.c_uint64(0)
hist = np.zeros(n_hash + 1, dtype=np.uint64)
if done.value:
    t_hist = wall(lambda: check(L.dyna_mh_plan_count_histogram(p, ptr(hist, C.c_uint64), None)))
    check(L.dyna_mh_plan_checksum(p, C.byref(h), None))
    print("  join: histogram %.2f ms, checksum %016x, nonzero pairs %d" % (t_hist, h.value, int(hist[1:].sum())))
    t_dens = wall(lambda: L.dyna_mh_plan_counts_device_ptr(p) or None, reps=1)
    hs = h.value
if n <= 120000:
    t_dense = wall(lambda: check(L.dyna_mh_plan_run_match(p, None)))
    check(L.dyna_mh_plan_checksum(p, C.byref(h), None))
    hist2 = np.zeros(n_hash + 1, dtype=np.uint64)
    check(L.dyna_mh_plan_count_histogram(p, ptr(hist2, C.c_uint64), None))
    print("  all-pairs kernel %.2f ms -> %.3g pairs/s, checksum %016x%s" % (t_dense, pairs / ((t_sig + t_dense) * 1e-3), h.value,
          (" (same: %s, histogram same: %s)" % (h.value == hs, bool((hist == hist2).all()))) if done.value else ""))
L.dyna_mh_plan_destroy(p)
