"""Multi-GPU layer of the KPopCount path: one process per GPU (torchrun), read-chunk sharding, table reduce.

The path shards naturally: a read stream is cut at record boundaries into one contiguous range per
rank, every rank fills its own dense 4^k table, and ONE exchange step follows -- the sum of the tables, an NCCL
all-reduce over NVLink (gloo on CPU in the tests).  Counts are integers, so the sum is exact and order-free; the
only care needed is counter width: the per-rank tables are u32 and their sum may not fit.
"""
import torch
import torch.distributed as dist

U32_LIMIT = 1 << 32


def shard_records(total_records, world_size, rank):
    """Contiguous, balanced record range [first, last) of `rank`; ranges tile [0, total) in rank order."""
    base, extra = divmod(total_records, world_size)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def can_reduce_as_u32(local_max, group=None):
    """True when no bin of the summed table can reach 2^32: the sum of the per-rank maxima bounds every bin."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return True
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([int(local_max)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, group=group)
    return int(t.item()) < U32_LIMIT


def reduce_dense_tables(lo_i32, promote, local_max, group=None):
    """Sum the dense tables of all ranks in place (every rank ends with the total).

    lo_i32   : the rank's u32 table viewed as int32 (two's complement addition is the same bit pattern as u32 addition)
    promote  : callable returning an int64 view of the table after folding lo into it (kpc_dense_promote + hi pointer);
               only called when a u32 sum could wrap
    returns the tensor that now holds the totals (lo_i32 or the int64 one)
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return lo_i32
    if can_reduce_as_u32(local_max, group):
        dist.all_reduce(lo_i32, group=group)
        return lo_i32
    hi_i64 = promote()
    dist.all_reduce(hi_i64, group=group)
    return hi_i64


class DeviceArray:
    """Zero-copy torch view of device memory owned by libkpopcount_gpu (via __cuda_array_interface__)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def table_views(counter):
    """(int32 view of the u32 table, promote()) for a KMerCounter on the dense path."""
    lo, _hi, nbins = counter.dense_table()
    lo_t = torch.as_tensor(DeviceArray(lo, nbins, "<i4"), device="cuda")

    def promote():
        counter.dense_promote()
        _lo, hi, n = counter.dense_table()
        return torch.as_tensor(DeviceArray(hi, n, "<i8"), device="cuda")

    return lo_t, promote
