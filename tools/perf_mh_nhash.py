import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo')
from dynaalign_b200 import synth
from dynaalign_b200._lib import check, flatten, lib, ptr
L = lib()
for n_hash in [int(a) for a in sys.argv[1:]] or (50, 100, 500):
    n = 32768
    seqs = synth.peptides_uniform(n)
    res, off = flatten(seqs)
    seeds = np.zeros(n_hash, dtype=np.uint32); check(L.dyna_hashfamily_seeds(42, n_hash, ptr(seeds, C.c_uint32)))
    p = L.dyna_mh_plan_create(n, n_hash, 0, n, 0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(L.dyna_mh_plan_upload_sequences(p, ptr(res, C.c_uint8), ptr(off, C.c_int64), 2 if n_hash == 50 else 4, ptr(seeds, C.c_uint32), st))
    check(L.dyna_mh_plan_run_signatures(p, st)); check(L.dyna_mh_plan_run_match(p, st)); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): check(L.dyna_mh_plan_run_match(p, st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("n_hash=%d n=%d match %.3f ms -> %.3e pairs/s" % (n_hash, n, ms, L.dyna_mh_plan_pairs(p) / ms * 1e3))
    L.dyna_mh_plan_destroy(p)
