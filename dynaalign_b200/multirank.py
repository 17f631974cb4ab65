"""One-process-per-GPU deployment of the MinHash half.

The pair space is sharded by row blocks with no data-path collective (`partition_rows`).  The one step every rank would
otherwise repeat in full is the exact 16-bit relabelling of the signature rows (segmented sort + rank scatter, ~3 ms
at 100,000 x 500): here each rank relabels `code_rows / world` packed rows and the table is completed by ONE
all-gather over NVLink (NCCL through torch.distributed), plus a max-reduce of the overflow flag that gates the
32-bit fallback kernel.  torch is plumbing only: the tensors below are zero-copy views of the plan's device buffers.
"""
import ctypes as C

from ._lib import check, lib


class _DeviceView:
    """Zero-copy handle on plan-owned device memory for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, shape, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def shard_bounds(code_rows, world, rank):
    """Packed code rows [begin, end) relabelled by `rank`; None when the exchange does not apply (single rank, no code
    table, or rows not divisible by the world size -> every rank relabels everything, as before)."""
    if world <= 1 or code_rows <= 0 or code_rows % world:
        return None
    per = code_rows // world
    return rank * per, (rank + 1) * per


def exchange_codes(table, overflow, world, rank, dist):
    """Complete the [code_rows, pitch] table in place from every rank's own rows; overflow <- max over ranks."""
    per = table.shape[0] // world
    dist.all_gather_into_tensor(table.reshape(-1), table[rank * per:(rank + 1) * per].reshape(-1))
    dist.all_reduce(overflow, op=dist.ReduceOp.MAX)


class ShardedSignatures:
    """run_signatures for one rank of `world`: K1 + layout transform locally, relabelling sharded + all-gathered."""

    def __init__(self, plan, world, rank, dist, torch, device):
        self.plan, self.world, self.rank, self.dist = plan, world, rank, dist
        L = lib()
        rows = L.dyna_mh_plan_code_rows(plan)
        self.bounds = shard_bounds(rows, world, rank)
        if self.bounds is not None:
            pitch = L.dyna_mh_plan_code_row_bytes(plan) // 4
            self.table = torch.as_tensor(_DeviceView(L.dyna_mh_plan_codes_device_ptr(plan), (rows, pitch)), device=device)
            self.overflow = torch.as_tensor(_DeviceView(L.dyna_mh_plan_overflow_device_ptr(plan), (1,)), device=device)

    @property
    def sharded(self):
        return self.bounds is not None

    def run(self, stream):
        """`stream` must be torch's current stream on this device, so the collective is ordered behind the kernels."""
        L = lib()
        if self.bounds is None:
            check(L.dyna_mh_plan_run_signatures(self.plan, stream))
            return
        check(L.dyna_mh_plan_run_signatures_shard(self.plan, self.bounds[0], self.bounds[1], stream))
        exchange_codes(self.table, self.overflow, self.world, self.rank, self.dist)
