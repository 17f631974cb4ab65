"""GPU, >= 2 devices: one process driving several GPUs through the host entry points (row blocks, one host thread
per device, host-side scatter) gives bit-identical matrices to the single-GPU run."""
import pytest

import dynaalign_b200 as da
from dynaalign_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _need2():
    if _lib.lib().dyna_device_count() < 2:
        pytest.skip("needs 2 CUDA devices")


def test_similarityNW_two_gpus_identical():
    _need2()
    seqs = [s.decode() for s in synth.proteins_families(260)]
    a = da.similarityNW(seqs, n_gpus=1)
    b = da.similarityNW(seqs, n_gpus=2)
    assert a.tobytes(order="F") == b.tobytes(order="F")


def test_similarityMH_two_gpus_identical():
    _need2()
    peps = [s.decode() for s in synth.peptides_clustered(3000, children=20)]
    a = da.similarityMH(peps, 4, 100, seed=42, n_gpus=1)
    b = da.similarityMH(peps, 4, 100, seed=42, n_gpus=2)
    assert a.tobytes(order="F") == b.tobytes(order="F")


def _nccl_worker(rank, world, port_no, q):
    import ctypes as C
    import os
    import sys

    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import dynaalign_b200 as da
    from dynaalign_b200 import _lib, synth
    from dynaalign_b200.multirank import ShardedSignatures
    L = _lib.lib()
    n, n_hash, k = 5000, 100, 4
    seqs = synth.peptides_clustered(n, children=20)
    res, off = _lib.flatten(seqs)
    seeds = da.hashfamily_seeds(42, n_hash)
    b = da.partition_rows(n, world)
    plan = L.dyna_mh_plan_create(n, n_hash, int(b[rank]), int(b[rank + 1]), rank)
    assert plan, _lib.last_error()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.dyna_mh_plan_upload_sequences(plan, _lib.ptr(res, C.c_uint8), _lib.ptr(off, C.c_int64), k, _lib.ptr(seeds, C.c_uint32), st))
    sh = ShardedSignatures(plan, world, rank, dist, torch, torch.device("cuda", rank))
    sh.run(st)
    _lib.check(L.dyna_mh_plan_run_match(plan, st))
    got = np.zeros(L.dyna_mh_plan_pairs(plan), dtype=np.uint16)
    _lib.check(L.dyna_mh_plan_fetch_counts(plan, _lib.ptr(got, C.c_uint16), st))
    L.dyna_mh_plan_destroy(plan)
    q.put((rank, sh.sharded, got))
    dist.barrier()
    dist.destroy_process_group()


def test_minhash_two_ranks_sharded_relabelling_over_nccl():
    # one process per GPU: row-block plans, relabelling sharded by code rows and completed by one NCCL all-gather
    _need2()
    import numpy as np
    import torch.multiprocessing as mp
    import os
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29700 + (os.getpid() % 90)
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    parts = sorted((q.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert parts[0][1] and parts[1][1]  # the exchange path was taken
    seqs = synth.peptides_clustered(5000, children=20)
    want = da.mh_match_counts(da.mh_signatures(seqs, 4, da.hashfamily_seeds(42, 100)))
    assert (np.concatenate([parts[0][2], parts[1][2]]) == want).all()
