"""GPU, >= 2 devices.  One process driving several GPUs through the R-facing entry points (row blocks, one host
thread per device, column blocks of the result gathered from the other devices' slabs by peer loads) and one process
per GPU over NCCL -- every result is compared with the ORACLE (the C port pinned to the compiled reference), not with
another GPU run."""
import numpy as np
import pytest

import dynaalign_b200 as da
from dynaalign_b200 import _lib, synth
from oracle import port

pytestmark = pytest.mark.gpu


def _need2():
    if _lib.lib().dyna_device_count() < 2:
        pytest.skip("needs 2 CUDA devices")


def test_similarityNW_two_gpus_equals_oracle():
    _need2()
    # 200 related proteins of ~330 aa plus the shapes that change kernel class (short, empty, > 384 rows)
    seqs = [s.decode() for s in synth.proteins_families(200)]
    rng = np.random.default_rng(5)
    al = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)
    seqs += [al[rng.integers(0, 20, size=int(L))].tobytes().decode() for L in (0, 7, 12, 31, 33, 400, 566, 700)]
    want = port.similarityNW(seqs)
    got2 = da.similarityNW(seqs, n_gpus=2)
    assert got2.tobytes(order="F") == want.tobytes(order="F")
    got1 = da.similarityNW(seqs, n_gpus=1)
    assert got1.tobytes(order="F") == want.tobytes(order="F")


def test_similarityMH_two_gpus_equals_oracle():
    _need2()
    peps = [s.decode() for s in synth.peptides_clustered(1500, children=20)] + ["", "AC", "ACD"]
    want = port.similarityMH(peps, 4, 100, 42)
    got = da.similarityMH(peps, 4, 100, seed=42, n_gpus=2)
    assert got.tobytes(order="F") == want.tobytes(order="F")


def test_minhash_r_distance_two_gpus_equals_oracle():
    _need2()
    # the R pipeline's distance matrix (1 - mean(==), long-double mean) from signatures, two devices
    rng = np.random.default_rng(3)
    sig = rng.integers(0, 6, size=(700, 37)).astype(np.uint32)
    want = port.mh_distance_matrix(sig)
    got = np.zeros((700, 700), dtype=np.float64, order="F")
    _lib.check(_lib.lib().dyna_mh_match_matrix(_lib.ptr(sig, __import__("ctypes").c_uint32), 700, 37, _lib.MH_DISTANCE,
                                              _lib.ptr(got, __import__("ctypes").c_double), 2))
    eq = (sig[:, None, :] == sig[None, :, :]).sum(axis=2)
    ref = 1.0 - (eq.astype(np.longdouble) / np.longdouble(37)).astype(np.float64)
    np.fill_diagonal(ref, 0.0)
    assert got.tobytes(order="F") == np.asfortranarray(ref).tobytes(order="F")
    assert got.tobytes(order="F") == np.asfortranarray(want).tobytes(order="F")


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
    want = port.mh_match_counts(port.mh_signatures(seqs, 4, port.hashfamily_seeds(42, 100)))  # the oracle, not a GPU run
    assert (np.concatenate([parts[0][2], parts[1][2]]) == want).all()
