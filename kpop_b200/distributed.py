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


def reduce_dense_tables(lo_i32, promote, local_max, group=None, has_hi=False):
    """Sum the dense tables of all ranks in place (every rank ends with the total).

    lo_i32   : the rank's u32 table viewed as int32 (two's complement addition is the same bit pattern as u32 addition)
    promote  : callable returning an int64 view of the table after folding lo into it (kpc_dense_promote + hi pointer);
               only called when a u32 sum could wrap
    has_hi   : this rank already keeps part of its counts in the 64-bit side table (kpc_dense_has_hi: a bin passed 2^31
               and was folded); then every rank must reduce in 64 bits, whatever the maxima say
    returns the tensor that now holds the totals (lo_i32 or the int64 one)
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return promote() if has_hi else lo_i32
    # the sum of the maxima bounds every bin of the sum; one extra unit per rank with a side table forces the wide path
    if can_reduce_as_u32(int(local_max) + (U32_LIMIT if has_hi else 0), group):
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


def _host_census(path, lo, hi, want_first=4, block=8 << 20):
    """(number of line feeds in [lo, hi), file offsets of the first `want_first` of them), reading bounded blocks."""
    import numpy as np
    count, first = 0, []
    with open(path, "rb") as f:
        pos = lo
        while pos < hi:
            f.seek(pos)
            b = np.frombuffer(f.read(min(block, hi - pos)), dtype=np.uint8)
            if b.size == 0:
                break
            if len(first) < want_first:
                nz = np.flatnonzero(b == 10)
                first += [pos + int(x) for x in nz[: want_first - len(first)]]
                count += int(nz.size)
            else:
                count += int(np.count_nonzero(b == 10))
            pos += b.size
    return count, first


def _all_gather_i64(value, rank, world_size, group):
    if world_size == 1:
        return [int(value)]
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.zeros(world_size, dtype=torch.int64, device=dev)
    t[rank] = int(value)
    dist.all_reduce(t, group=group)  # a sum of one-hot vectors: an all-gather of 8 bytes per rank
    return [int(x) for x in t.tolist()]


def record_aligned_ranges(size, lo, hi, n_newlines, first_newlines, byte_before_lo, rank, world_size, group=None):
    """The exchange step of the sharding: from every rank's line-feed count (and the positions of its first four line
    feeds) to the record-aligned byte range [start, end) of this rank.  The ranges tile [0, size)."""
    counts = _all_gather_i64(n_newlines, rank, world_size, group)
    lines_before = sum(counts[:rank])          # line feeds before `lo` = index of the line `lo` lies in
    start = -1
    if rank == 0:
        start = 0
    elif lo > 0 and byte_before_lo == 10 and lines_before % 4 == 0:
        start = lo                             # `lo` itself is the first byte of a record
    else:
        k = (-lines_before) % 4 or 4           # the k-th line feed of the range is followed by the start of a record
        if len(first_newlines) >= k:
            start = first_newlines[k - 1] + 1
    starts = [x - 1 for x in _all_gather_i64(start + 1, rank, world_size, group)]
    nxt = size
    ends = [0] * world_size
    for r in range(world_size - 1, -1, -1):    # ranks without a boundary in their range take nothing
        if starts[r] < 0:
            starts[r] = nxt
        ends[r] = nxt
        nxt = starts[r]
    return starts[rank], ends[rank]


def shard_fastq_byte_range(path, group=None, rank=None, world_size=None):
    """Byte range [start, end) of `path` that this rank counts: contiguous, record aligned, the ranges tile the file.

    FASTQ records are four LINES (Files.ml:201-221), and a line that starts with '@' may just as well be a quality line,
    so a record boundary cannot be recognised locally.  Exact recipe: every rank counts the line feeds of its tentative
    range, the counts are all-gathered (8 bytes per rank -- the only cross-shard data), their prefix sum gives the index of
    the line each range starts in, and the rank moves its start forward to the first line whose index is a multiple of 4.
    The last rank keeps the tail of the file whatever it holds (an incomplete last record is dropped by the counter, as
    FASTQ.iter_se does).  This entry point takes the census on the host (bounded blocks); count_fastq_sharded takes it on
    the device while the shard is being uploaded.
    """
    import os
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    size = os.path.getsize(path)
    lo, hi = _tentative_range(size, world_size, rank)
    n, first = _host_census(path, lo, hi)
    before = 10
    if lo > 0:
        with open(path, "rb") as f:
            f.seek(lo - 1)
            before = f.read(1)[0]
    return record_aligned_ranges(size, lo, hi, n, first, before, rank, world_size, group)


def _raise_together(err, rank, world_size, group):
    """A rank that failed must not leave the others waiting in a collective: the failure is shared first."""
    flags = _all_gather_i64(1 if err else 0, rank, world_size, group)
    if any(flags):
        if err:
            raise err
        raise RuntimeError("rank(s) %s failed while counting their shard" % [r for r, f in enumerate(flags) if f])


def count_fastq_sharded(path, k=12, label="sample", device=None, group=None, chunk_bytes=64 << 20, lookahead=1 << 20):
    """KPopCount -k K -l LABEL -s PATH on all the GPUs of the process group (read-chunk sharding).

    Every rank uploads its tentative byte range of the file through two pinned staging buffers (reads overlap the copies),
    takes the line-feed census of the shard ON THE DEVICE (kpc_count_newlines), learns from the 8-byte-per-rank exchange where
    its first record starts, counts its record-aligned range straight from device memory (kpc_feed_device) into its own
    dense 4^k table, the tables are summed (NCCL all-reduce over NVLink) and rank 0 formats.  Returns the spectrum text on
    rank 0 and None elsewhere.  Dense-table configurations only (k <= 12 for DNA)."""
    import os
    import numpy as np
    from .counter import KMerCounter
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if device is None:
        device = torch.cuda.current_device()
    size = os.path.getsize(path)
    lo, hi = _tentative_range(size, world, rank)
    top = min(size, hi + lookahead)            # the shard ends where the next one starts: a little past `hi`
    text, err = None, None
    with KMerCounter(k=k, label=label, device=device) as kc:
        lib_stream = torch.cuda.ExternalStream(kc.stream_handle())
        dev = torch.empty(top - lo + 64, dtype=torch.uint8, device="cuda")
        first, before, n_nl = [], 10, 0
        try:
            # ---- upload [lo, top) through two pinned buffers; the first four line feeds are looked up on the way ----
            stage = [torch.empty(chunk_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
            done = [torch.cuda.Event(), torch.cuda.Event()]
            copy = torch.cuda.Stream()
            with open(path, "rb", buffering=0) as f:
                if lo > 0:
                    f.seek(lo - 1)
                    before = f.read(1)[0]
                f.seek(lo)
                pos, i = lo, 0
                while pos < top:
                    b = stage[i & 1]
                    if i >= 2:
                        done[i & 1].synchronize()
                    n = f.readinto(memoryview(b.numpy())[: min(chunk_bytes, top - pos)])
                    if not n:
                        break
                    if len(first) < 4 and pos < hi:
                        nz = np.flatnonzero(b.numpy()[: min(n, hi - pos)] == 10)
                        first += [pos + int(x) for x in nz[: 4 - len(first)]]
                    with torch.cuda.stream(copy):
                        dev[pos - lo: pos - lo + n].copy_(b[:n], non_blocking=True)
                        done[i & 1].record(copy)
                    pos += n
                    i += 1
            copy.synchronize()
            # ---- census on the device, exchange, record-aligned range ----
            n_nl = kc.count_newlines(dev.data_ptr(), hi - lo)
        except Exception as e:  # noqa: BLE001 -- shared with the other ranks below
            err = e
        _raise_together(err, rank, world, group)
        start, end = record_aligned_ranges(size, lo, hi, n_nl, first, before, rank, world, group)
        try:
            if end > top:
                raise RuntimeError("a FASTQ record longer than the look-ahead (%d bytes) straddles two shards" % lookahead)
            kc.begin("single-end")
            if end > start:
                off, n = start - lo, end - start
                src = dev
                if off % 16:                   # kpc_feed_device wants a 16-byte aligned pointer: one device-to-device move
                    src = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
                    src[:n].copy_(dev[off: off + n])
                    off = 0
                torch.cuda.synchronize()
                kc.feed_device(src.data_ptr() + off, n, eof=True)
            else:
                kc.feed(b"", eof=True)
            kc.end()                           # raises on a malformed record of this shard
        except Exception as e:  # noqa: BLE001
            err = e
        _raise_together(err, rank, world, group)
        lo_t, promote = table_views(kc)
        torch.cuda.current_stream().wait_stream(lib_stream)
        reduce_dense_tables(lo_t, promote, kc.dense_max(), group, has_hi=kc.dense_has_hi())
        torch.cuda.synchronize()
        if rank == 0:
            kc.finish()
            text = kc.take_text()
    return text
