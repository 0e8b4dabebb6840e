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


# ------------------------------------------------------------------------------------------------------------
# read-chunk sharding of ONE FASTQ file (configuration C5)
# ------------------------------------------------------------------------------------------------------------
def _tentative_range(size, world_size, rank):
    return size * rank // world_size, size * (rank + 1) // world_size


def shard_fastq_byte_range(path, group=None, rank=None, world_size=None):
    """Byte range [start, end) of `path` that this rank counts: contiguous, record aligned, the ranges tile the file.

    FASTQ records are four LINES (Files.ml:201-221), and a line that starts with '@' may just as well be a quality line,
    so a record boundary cannot be recognised locally.  Exact recipe: every rank counts the line feeds of its tentative
    range, the counts are all-gathered (8 bytes per rank -- the only cross-shard data), their prefix sum gives the index of
    the line each range starts in, and the rank moves its start forward to the first line whose index is a multiple of 4.
    The last rank keeps the tail of the file whatever it holds (an incomplete last record is dropped by the counter, as
    FASTQ.iter_se does).
    """
    import numpy as np
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    size = __import__("os").path.getsize(path)
    lo, hi = _tentative_range(size, world_size, rank)
    buf = np.memmap(path, dtype=np.uint8, mode="r") if size else np.zeros(0, dtype=np.uint8)
    nl = np.flatnonzero(buf[lo:hi] == 10) if hi > lo else np.zeros(0, dtype=np.int64)
    counts = [0] * world_size
    if world_size > 1:
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.zeros(world_size, dtype=torch.int64, device=dev)
        t[rank] = int(nl.size)
        dist.all_reduce(t, group=group)
        counts = [int(x) for x in t.tolist()]
    lines_before = sum(counts[:rank])          # line feeds before `lo` = index of the line `lo` lies in
    start = None
    if rank == 0:
        start = 0
    else:
        # the k-th line feed of the range (1-based) is followed by the start of line lines_before + k
        k = (-lines_before) % 4 or 4
        # ... unless `lo` itself is a line start of the right index (the byte before it is a line feed)
        if lo > 0 and buf[lo - 1] == 10 and lines_before % 4 == 0:
            start = lo
        elif nl.size >= k:
            start = lo + int(nl[k - 1]) + 1
    # ranks without a boundary in their range take nothing: their start is the next rank's start
    starts = [None] * world_size
    if world_size > 1:
        t = torch.full((world_size,), -1, dtype=torch.int64, device=dev)
        t[rank] = -1 if start is None else int(start)
        u = torch.zeros(world_size, dtype=torch.int64, device=dev)
        u[rank] = t[rank] + 1                   # all_reduce(sum) of (start + 1), zeros elsewhere
        dist.all_reduce(u, group=group)
        starts = [int(x) - 1 for x in u.tolist()]
    else:
        starts = [0]
    nxt = size
    ends = [0] * world_size
    for r in range(world_size - 1, -1, -1):
        if starts[r] < 0:
            starts[r] = nxt
        ends[r] = nxt
        nxt = starts[r]
    return starts[rank], ends[rank]


def count_fastq_sharded(path, k=12, label="sample", device=None, group=None, chunk_bytes=1 << 28):
    """KPopCount -k K -l LABEL -s PATH on all the GPUs of the process group: every rank counts its record-aligned byte
    range of the file into its own dense 4^k table, the tables are summed (NCCL all-reduce over NVLink), rank 0 formats.
    Returns the spectrum text on rank 0 and None elsewhere.  Dense-table configurations only (k <= 12 for DNA)."""
    import numpy as np
    from .counter import KMerCounter
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    start, end = shard_fastq_byte_range(path, group)
    if device is None:
        device = torch.cuda.current_device()
    text = None
    with KMerCounter(k=k, label=label, device=device) as kc:
        kc.begin("single-end")
        if end > start:
            buf = np.memmap(path, dtype=np.uint8, mode="r")
            pos = start
            while pos < end:
                n = min(chunk_bytes, end - pos)
                kc.feed_pointer(buf.ctypes.data + pos, n, eof=(pos + n >= end))  # straight out of the page cache
                pos += n
        else:
            kc.feed(b"", eof=True)
        kc.end()
        lo_t, promote = table_views(kc)
        torch.cuda.current_stream().wait_stream(torch.cuda.ExternalStream(kc.stream_handle()))
        reduce_dense_tables(lo_t, promote, kc.dense_max(), group)
        torch.cuda.synchronize()
        if rank == 0:
            kc.finish()
            text = kc.take_text()
    return text
