"""CPU, world_size 2 over gloo: the N>1 path of bench.py / the multi-process deployment.  Each rank takes the row
block dyna_partition_rows assigns it, produces its slab (here with the oracle, since there is no GPU on this box),
and the slabs must tile the packed triangle exactly -- no overlap, no gap, no data-path collective needed."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port_no, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dynaalign_b200 as da
    from oracle import port

    rng = np.random.default_rng(3)  # same data on every rank
    al = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)
    seqs = [al[rng.integers(0, 20, size=int(rng.integers(0, 60)))].tobytes().decode() for _ in range(41)]
    n = len(seqs)
    # NW: blocks balanced by cells, diagonal included
    b = da.partition_rows(n, world, weights=[len(s) for s in seqs], include_diagonal=True)
    mt, ln = port.nw_pair_stats(seqs, row_begin=int(b[rank]), row_end=int(b[rank + 1]))
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(mt)], dtype=torch.int64))
    # MinHash: blocks balanced by pairs, strict triangle
    sig = port.mh_signatures(seqs, 3, port.hashfamily_seeds(5, 40))
    mb = da.partition_rows(n, world)
    cnt = port.mh_match_counts(sig, int(mb[rank]), int(mb[rank + 1]))
    # the verification gathers slabs (test only; the product path has no collective)
    maxlen = n * (n + 1) // 2
    pad = torch.zeros(maxlen, dtype=torch.int64)
    pad[:len(mt)] = torch.from_numpy(mt.astype(np.int64) * 100000 + ln.astype(np.int64))
    allnw = [torch.zeros(maxlen, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allnw, pad)
    padc = torch.zeros(maxlen, dtype=torch.int64)
    padc[:len(cnt)] = torch.from_numpy(cnt.astype(np.int64))
    allmh = [torch.zeros(maxlen, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allmh, padc)
    csz = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(csz, torch.tensor([len(cnt)], dtype=torch.int64))
    if rank == 0:
        full_m, full_l = port.nw_pair_stats(seqs)
        got = np.concatenate([allnw[r][:int(sizes[r])].numpy() for r in range(world)])
        ok_nw = len(got) == len(full_m) and (got == full_m.astype(np.int64) * 100000 + full_l.astype(np.int64)).all()
        gotc = np.concatenate([allmh[r][:int(csz[r])].numpy() for r in range(world)])
        full_c = port.mh_match_counts(sig)
        ok_mh = len(gotc) == len(full_c) and (gotc == full_c.astype(np.int64)).all()
        q.put((bool(ok_nw), bool(ok_mh), b.tolist(), mb.tolist()))
    dist.destroy_process_group()


def test_two_ranks_tile_the_triangle():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok_nw, ok_mh, b, mb = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok_nw and ok_mh
    assert b[0] == 0 and b[-1] == 41 and mb[0] == 0 and mb[-1] == 41 and 0 < b[1] < 41 and 0 < mb[1] < 41


def _codes_worker(rank, world, port_no, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dynaalign_b200.multirank import exchange_codes, shard_bounds

    rows, pitch = 16, 384
    b = shard_bounds(rows, world, rank)
    table = torch.full((rows, pitch), -1, dtype=torch.int32)
    table[b[0]:b[1]] = (torch.arange(b[0], b[1], dtype=torch.int32)[:, None] * 1000 + torch.arange(pitch, dtype=torch.int32)[None, :])
    overflow = torch.tensor([1 if rank == 1 else 0], dtype=torch.int32)
    exchange_codes(table, overflow, world, rank, dist)
    want = torch.arange(rows, dtype=torch.int32)[:, None] * 1000 + torch.arange(pitch, dtype=torch.int32)[None, :]
    q.put((rank, bool((table == want).all()), int(overflow[0]), b))
    dist.destroy_process_group()


def test_code_table_exchange_two_ranks():
    # the one exchange step of the multi-rank MinHash path: each rank owns code_rows / world rows of the relabelled
    # table, an in-place all-gather completes it and the overflow gate becomes the max over ranks
    from dynaalign_b200.multirank import shard_bounds
    assert shard_bounds(256, 1, 0) is None and shard_bounds(0, 2, 0) is None and shard_bounds(250, 4, 1) is None
    assert [shard_bounds(256, 8, r) for r in (0, 7)] == [(0, 32), (224, 256)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_codes_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [(0, True, 1, (0, 8)), (1, True, 1, (8, 16))]


def _checksum_worker(rank, world, port_no, q, shift):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dynaalign_b200 as da
    from oracle import port

    rng = np.random.default_rng(11)
    al = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)
    seqs = [al[rng.integers(0, 20, size=int(rng.integers(1, 40)))].tobytes().decode() for _ in range(53)]
    n = len(seqs)
    b = da.partition_rows(n, world, weights=[len(s) for s in seqs], include_diagonal=True).tolist()
    if shift and rank == 1:
        b[1] += shift  # a deliberate off-by-one in the partition: rank 1 starts one row late
    rb, re_ = int(b[rank]), int(b[rank + 1])
    mt, ln = port.nw_pair_stats(seqs, row_begin=rb, row_end=re_)
    first = rb * n - rb * (rb - 1) // 2
    mine = (da.checksum(mt, first), da.checksum(ln, first))
    # bench.py's parity step: gather every rank's pair of sums over the gloo group, add mod 2^64
    out = [None] * world
    dist.all_gather_object(out, mine)
    total = tuple(sum(x[k] for x in out) & ((1 << 64) - 1) for k in range(2))
    if rank == 0:
        wm, wl = port.nw_pair_stats(seqs)
        q.put(total == (da.checksum(wm), da.checksum(wl)))
    dist.destroy_process_group()


@pytest.mark.parametrize("shift,expect", [(0, True), (1, False)])
def test_slab_checksums_add_up_only_for_an_exact_partition(shift, expect):
    # the multi-GPU parity check of bench.py on CPU: position-weighted checksums of the ranks' slabs sum to the
    # single-process value, and an off-by-one row in the partition turns the check red
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 30100 + (os.getpid() % 300) + shift
    procs = [ctx.Process(target=_checksum_worker, args=(r, 2, port_no, q, shift)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got is expect
