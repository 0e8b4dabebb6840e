"""Multi-GPU layer of the KPopCount path: one process per GPU (torchrun), read-chunk sharding, table reduce.

The path shards naturally: a read stream is cut at record boundaries into one contiguous range per
rank, every rank fills its own dense 4^k table, and ONE exchange step follows -- the sum of the tables, an NCCL
all-reduce over NVLink (gloo on CPU in the tests).  Counts are integers, so the sum is exact and order-free; the
only care needed is counter width: the per-rank tables are u32 and their sum may not fit.
"""
import torch
import torch.distributed as dist

from . import _native as N

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
    # index of the record that starts at `start`: the line feeds before it, over four
    first_line = 0
    if rank > 0 and start >= 0:
        first_line = lines_before if start == lo else lines_before + ((-lines_before) % 4 or 4)
    starts = [x - 1 for x in _all_gather_i64(start + 1, rank, world_size, group)]
    first_lines = _all_gather_i64(first_line, rank, world_size, group)
    total_lines = sum(counts)
    nxt, nxt_line = size, total_lines
    ends = [0] * world_size
    for r in range(world_size - 1, -1, -1):    # ranks without a boundary in their range take nothing
        if starts[r] < 0:
            starts[r], first_lines[r] = nxt, nxt_line
        ends[r] = nxt
        nxt, nxt_line = starts[r], first_lines[r]
    record_aligned_ranges.first_record = first_lines[rank] // 4   # (side channel for count_fastq_sharded_sparse)
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


# ------------------------------------------------------------------------------------------------------------
# read-chunk sharding of a PAIR of FASTQ files ("both mate files must be cut at the same record index")
# ------------------------------------------------------------------------------------------------------------
def _default_device(device):
    """The GPU of this rank when the caller names none: torch's current device (torchrun scripts set it per rank)."""
    if device is not None:
        return device
    return torch.cuda.current_device() if torch.cuda.is_available() else 0


def _nth_line_feed(path, lo, hi, n, block=8 << 20):
    """File offset of the n-th (1-based) line feed of [lo, hi), or -1."""
    import numpy as np
    with open(path, "rb") as f:
        pos = lo
        while pos < hi:
            f.seek(pos)
            b = np.frombuffer(f.read(min(block, hi - pos)), dtype=np.uint8)
            if b.size == 0:
                break
            nz = np.flatnonzero(b == 10)
            if n <= nz.size:
                return pos + int(nz[n - 1])
            n -= int(nz.size)
            pos += b.size
    return -1


def _all_reduce_max_i64(values, world_size, group):
    if world_size == 1:
        return [int(v) for v in values]
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [int(x) for x in t.tolist()]


def pair_aligned_ranges(path1, path2, group=None, rank=None, world_size=None):
    """((start1, end1), (start2, end2), first_pair, n_pairs): the byte ranges of the two mate files that hold the SAME
    pairs [first_pair, first_pair + n_pairs) for this rank; over the ranks the pairs tile [0, P), P = the number of
    complete records of the shorter file (FASTQ.iter_pe stops there, Files.ml:228-247; what lies behind is never read).

    Two exchanges of a few integers per rank: (1) every rank counts the line feeds of its tentative byte range of both
    files -> total lines -> P and the pair every rank starts with; (2) the rank whose tentative range holds the line feed
    that ends line 4 * boundary - 1 of a file looks its offset up; a MAX all-reduce hands every offset to everybody."""
    import os
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    paths = (path1, path2)
    sizes = [os.path.getsize(p) for p in paths]
    tent = [_tentative_range(sizes[f], world_size, rank) for f in (0, 1)]
    mine = [_host_census(paths[f], tent[f][0], tent[f][1], want_first=0)[0] for f in (0, 1)]
    counts = [_all_gather_i64(mine[f], rank, world_size, group) for f in (0, 1)]
    records = []
    for f in (0, 1):
        lines = sum(counts[f])
        if sizes[f]:
            with open(paths[f], "rb") as fh:
                fh.seek(sizes[f] - 1)
                if fh.read(1) != b"\n":
                    lines += 1                 # input_line returns a last line without line feed like any other
        records.append(lines // 4)             # an incomplete last record is dropped (Files.ml:217)
    n_total = min(records)
    bounds = [shard_records(n_total, world_size, r)[0] for r in range(world_size)] + [n_total]
    # offset of the first byte of record b of file f = one past the line feed number 4 b (1-based)
    want = []
    for f in (0, 1):
        before = sum(counts[f][:rank])         # line feeds before this rank's tentative range
        for b in bounds:
            off = -1
            if b == 0:
                off = 0
            else:
                nth = 4 * b
                if before < nth <= before + mine[f]:
                    at = _nth_line_feed(paths[f], tent[f][0], tent[f][1], nth - before)
                    if at < 0:
                        raise RuntimeError("%s changed while it was being cut" % paths[f])
                    off = at + 1
                elif nth > sum(counts[f]) and rank == world_size - 1:
                    off = sizes[f]             # the last record of the file ends without a line feed
            want.append(off)
    offs = _all_reduce_max_i64(want, world_size, group)
    n = world_size + 1
    r1 = (offs[rank], offs[rank + 1])
    r2 = (offs[n + rank], offs[n + rank + 1])
    if min(r1 + r2) < 0:
        raise RuntimeError("internal: a pair boundary was not located")
    return r1, r2, bounds[rank], bounds[rank + 1] - bounds[rank]


def count_fastq_pair_sharded(path1, path2, k=12, label="sample", device=None, group=None, lib=None, chunk_bytes=64 << 20):
    """KPopCount -k K -l LABEL -p PATH1 PATH2 on all the ranks of the process group, dense-table configurations: every rank
    counts the same range of PAIRS out of both files (pair_aligned_ranges), the tables are summed, rank 0 formats.  In the
    dense table the order of the mates does not matter (sums commute); what has to be exact is which records belong to a
    pair at all.  Returns the text on rank 0, None elsewhere."""
    from .counter import KMerCounter
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    device = _default_device(device)
    (s1, e1), (s2, e2), first_pair, _n = pair_aligned_ranges(path1, path2, group, rank, world)
    text, err = None, None
    with KMerCounter(k=k, label=label, device=device, lib=lib) as kc:
        try:
            kc.begin("paired-end")
            with open(path1, "rb") as f1, open(path2, "rb") as f2:
                pos, end, fh = [s1, s2], [e1, e2], [f1, f2]
                done = [False, False]
                while not all(done):
                    for m in (0, 1):
                        if done[m]:
                            continue
                        fh[m].seek(pos[m])
                        data = fh[m].read(min(chunk_bytes, end[m] - pos[m]))
                        pos[m] += len(data)
                        done[m] = pos[m] >= end[m]
                        kc.feed(data, mate=m, eof=done[m])
            kc.end()
        except Exception as e:  # noqa: BLE001 -- shared with the other ranks below
            err = e
        # a malformed pair: the reference stops at the FIRST one and names the last of its eight lines, counted over both
        # files (Files.ml:241-243); a shard reports lines of its own range, so the number is moved and the smallest wins
        bad_line = 0
        if err is not None and getattr(err, "code", None) == N.KPC_E_MALFORMED_FASTQ:
            import re
            m = re.search(r"On line (\d+):", err.message)
            if m:
                bad_line = int(m.group(1)) + 8 * first_pair
                err = None
        lines = _all_gather_i64(bad_line, rank, world, group)
        _raise_together(err, rank, world, group)
        if any(lines):
            from .counter import KPopCountError
            raise KPopCountError(N.KPC_E_MALFORMED_FASTQ, "On line %d: Malformed FASTQ file" % min(x for x in lines if x))
        lo, _hi, nbins = kc.dense_table()
        if kc.backend() == "cuda":
            lo_t, promote = table_views(kc)
        else:                                  # test-only emulation: the table is host memory
            import ctypes
            import numpy as np
            lo_t = torch.from_numpy(np.ctypeslib.as_array(ctypes.cast(lo, ctypes.POINTER(ctypes.c_int32)), shape=(nbins,)))

            def promote():
                kc.dense_promote()
                _lo, hi, n = kc.dense_table()
                return torch.from_numpy(np.ctypeslib.as_array(ctypes.cast(hi, ctypes.POINTER(ctypes.c_int64)), shape=(n,)))
        kc.sync()
        reduce_dense_tables(lo_t, promote, kc.dense_max(), group, has_hi=kc.dense_has_hi())
        if kc.backend() == "cuda":
            torch.cuda.synchronize()
        if rank == 0:
            kc.finish()
            text = kc.take_text()
    return text


# ------------------------------------------------------------------------------------------------------------
# sparse merge: ONE large-k sample on all the GPUs (last bullet)
# ------------------------------------------------------------------------------------------------------------
def _u64_view(counter, ptr, n):
    """int64 torch view of n u64 words the library owns (device memory; host memory under the test-only emulation)."""
    if n == 0 or not ptr:
        return torch.zeros(0, dtype=torch.int64, device="cuda" if counter.backend() == "cuda" else "cpu")
    if counter.backend() == "cuda":
        return torch.as_tensor(DeviceArray(ptr, n, "<i8"), device="cuda")
    import ctypes
    import numpy as np
    return torch.from_numpy(np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_int64)), shape=(n,)))


def bucket_owner(keys_i64, n_buckets, world_size):
    """Owner rank of every key: OCaml bucket indices (key mod B) are cut into world_size contiguous ranges, so that the
    dumps of the owners, concatenated in rank order, are in Hashtbl.iter order."""
    return ((keys_i64 & (n_buckets - 1)) * world_size) // n_buckets


def exchange_entries(keys, counts, ranks, n_buckets, rank, world_size, group=None):
    """Every entry travels to the rank that owns its bucket (all-to-all with the sizes exchanged first)."""
    if world_size == 1:
        return keys, counts, ranks
    owner = bucket_owner(keys, n_buckets, world_size)
    order = torch.argsort(owner, stable=True)
    send_sizes = torch.bincount(owner, minlength=world_size).to(torch.int64)
    recv_sizes = torch.zeros_like(send_sizes)
    dist.all_to_all_single(recv_sizes, send_sizes, group=group)
    ss, rs = [int(x) for x in send_sizes.tolist()], [int(x) for x in recv_sizes.tolist()]
    out = []
    for t in (keys, counts, ranks):
        src = t[order].contiguous()
        dst = torch.empty(sum(rs), dtype=torch.int64, device=t.device)
        if dist.get_backend(group) == "gloo":   # gloo has no all_to_all_single with uneven splits on every version: send / recv
            reqs = []
            so = ro = 0
            for r in range(world_size):
                if r == rank:
                    dst[ro: ro + rs[r]] = src[so: so + ss[r]]
                else:
                    if ss[r]:
                        reqs.append(dist.isend(src[so: so + ss[r]].clone(), r, group=group))
                    if rs[r]:
                        reqs.append(dist.irecv(dst[ro: ro + rs[r]], r, group=group))
                so += ss[r]
                ro += rs[r]
            for q in reqs:
                q.wait()
        else:
            dist.all_to_all_single(dst, src, rs, ss, group=group)
        out.append(dst)
    return tuple(out)


def count_fastq_sharded_sparse(path, k=21, label="sample", device=0, group=None, max_results_size=16777216, lib=None,
                               chunk_bytes=32 << 20):
    """KPopCount -k K -l LABEL -s PATH for k > 12 (hash-table path) on all the ranks of the process group.

    Every rank counts its record-aligned byte range of the file with insertion ranks that are valid for the whole file,
    exports its distinct k-mers, the entries travel to the rank that owns their OCaml bucket, the owners merge them on the
    device (counts add, the smallest insertion rank wins: kpc_hash_import) and dump their bucket range; rank 0 returns the
    concatenation -- the bytes one KPopCount process prints -- and the others None.  Raises when the merged table could
    reach -M (the reference would then dump in the middle of the stream, which does not shard)."""
    from .counter import KMerCounter, KPopCountError
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    start, end = shard_fastq_byte_range(path, group)
    first_record = record_aligned_ranges.first_record
    err, text = None, None
    with KMerCounter(k=k, label=label, max_results_size=max_results_size, device=device, lib=lib) as kc:
        n_slots = 0
        try:
            kc.set_record_base(first_record)
            kc.begin("single-end")
            with open(path, "rb") as f:
                f.seek(start)
                pos = start
                if end == start:
                    kc.feed(b"", eof=True)
                while pos < end:
                    b = f.read(min(chunk_bytes, end - pos))
                    pos += len(b)
                    kc.feed(b, eof=(pos >= end))
            kc.end()
            kp, cp, rp, n_slots = kc.hash_export()
        except Exception as e:  # noqa: BLE001 -- shared with the other ranks below
            err = e
        _raise_together(err, rank, world, group)
        keys, counts, ranks = (_u64_view(kc, p, n_slots) for p in (kp, cp, rp))
        live = keys != -1                       # empty slots carry the key ~0
        keys, counts, ranks = keys[live], counts[live], ranks[live]
        total = sum(_all_gather_i64(int(keys.numel()), rank, world, group))
        if total >= max_results_size:
            raise KPopCountError(-9, "the merged table could reach -M: the reference would dump in the middle of the stream")
        keys, counts, ranks = exchange_entries(keys, counts, ranks, kc.bucket_count(), rank, world, group)
        keys, counts, ranks = keys.contiguous(), counts.contiguous(), ranks.contiguous()
        if keys.is_cuda:
            torch.cuda.synchronize()
        kc.take_text()
        kc.hash_import(keys.data_ptr(), counts.data_ptr(), ranks.data_ptr(), int(keys.numel()), clear_first=True)
        kc.finish()
        mine = kc.take_text()
    header = ("\t%s\n" % label).encode()
    if mine.startswith(header):
        mine = mine[len(header):]
    if world == 1:
        return header + mine
    parts = [None] * world
    dist.all_gather_object(parts, mine, group=group)
    return header + b"".join(parts) if rank == 0 else None


# ------------------------------------------------------------------------------------------------------------
# -L (one spectrum per read) on all the ranks ("-L mode shards by records")
# ------------------------------------------------------------------------------------------------------------
def _longest_line(path, lo, hi, block=8 << 20):
    """Length of the longest line that has a byte in [lo, hi) (pieces cut by the range ends count as they are)."""
    import numpy as np
    longest, run = 0, 0
    with open(path, "rb") as f:
        pos = lo
        while pos < hi:
            f.seek(pos)
            b = np.frombuffer(f.read(min(block, hi - pos)), dtype=np.uint8)
            if b.size == 0:
                break
            nz = np.flatnonzero(b == 10)
            if nz.size == 0:
                run += int(b.size)
            else:
                longest = max(longest, run + int(nz[0]))
                if nz.size > 1:
                    longest = max(longest, int(np.max(np.diff(nz))) - 1)
                run = int(b.size) - int(nz[-1]) - 1
            pos += b.size
    return max(longest, run)


def count_fastq_sharded_per_record(path, k=12, content=None, device=None, group=None, max_results_size=16777216, lib=None,
                                   chunk_bytes=32 << 20):
    """KPopCount -k K -L -s PATH on all the ranks of the process group: every rank prints the spectra of the records in
    its record-aligned byte range, rank 0 returns the concatenation in rank order (= record order), the others None.

    A record's spectrum depends on nothing but the record -- except through the bucket count of the reference's table,
    which only ever grows (when one record holds more than twice as many distinct k-mers as there are buckets) and then
    changes the order of every later spectrum.  That cannot happen when 2 * buckets >= the longest line of the file, which
    is checked here (the default -M gives 2^24 buckets); otherwise the call refuses instead of guessing."""
    from .counter import Content, KMerCounter, KPopCountError
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    device = _default_device(device)
    start, end = shard_fastq_byte_range(path, group, rank, world)
    buckets = 16
    while buckets < max_results_size:
        buckets *= 2
    longest = max(_all_gather_i64(_longest_line(path, start, end), rank, world, group))
    if longest > 2 * buckets:
        raise KPopCountError(-9, "a record of %d bytes could make the table of %d buckets grow: its later spectra would "
                                 "depend on earlier shards" % (longest, buckets))
    err, mine, bad_line = None, b"", 0
    with KMerCounter(k=k, content=Content.DNA_ds if content is None else content, label="", device=device, lib=lib,
                     max_results_size=max_results_size) as kc:
        try:
            kc.begin("single-end")
            with open(path, "rb") as f:
                f.seek(start)
                pos = start
                if end == start:
                    kc.feed(b"", eof=True)
                while pos < end:
                    b = f.read(min(chunk_bytes, end - pos))
                    pos += len(b)
                    kc.feed(b, eof=(pos >= end))
            kc.end()
            kc.finish()
        except Exception as e:  # noqa: BLE001 -- shared with the other ranks below
            err = e
        mine = kc.take_text() if err is None or getattr(err, "code", None) == N.KPC_E_MALFORMED_FASTQ else b""
    # a malformed record: the reference prints every spectrum before it, then fails with its line number (Files.ml:213)
    if err is not None and getattr(err, "code", None) == N.KPC_E_MALFORMED_FASTQ:
        import re
        m = re.search(r"On line (\d+):", err.message)
        if m:
            bad_line = int(m.group(1)) + 4 * record_aligned_ranges.first_record
            err = None
    lines = _all_gather_i64(bad_line, rank, world, group)
    _raise_together(err, rank, world, group)
    parts = [mine]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine, group=group)
    if any(lines):
        first_bad = min(r for r in range(world) if lines[r])
        e = KPopCountError(N.KPC_E_MALFORMED_FASTQ, "On line %d: Malformed FASTQ file" % lines[first_bad])
        e.partial_text = b"".join(parts[: first_bad + 1])   # what the reference had printed when it failed
        raise e
    return b"".join(parts) if rank == 0 else None
