"""world_size = 2 on CPU (gloo): the host logic of the multi-GPU path -- record sharding and the exact table reduce,
including the switch to a 64-bit reduction when a u32 sum could wrap.  The per-shard tables come from the CPU checker
(oracle/fast_dense.cpp); on the GPU box the same functions are driven by bench.py with the device tables."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ORACLE_DIR, ROOT

K = 8
N_REC = 4001


def _worker(rank, world, port, tmp, big):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from kpop_b200.distributed import reduce_dense_tables, shard_records
    from test_oracle_fastdense import SYNTH, dense_table, fastdense
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, last = shard_records(N_REC, world, rank)
    data = subprocess.run([SYNTH, str(first), str(last - first), "3"], stdout=subprocess.PIPE, check=True).stdout
    lib = fastdense()
    t = dense_table(lib, data, K, threads=2)
    if big == 1:  # pretend one bin is close to wrapping on every rank
        t[5] = np.uint32(0xF0000000)
    lo = torch.from_numpy(t.view(np.int32).copy())

    def promote():
        return torch.from_numpy(t.astype(np.int64))

    # big == 2: no bin is large, but rank 1 already keeps counts in its 64-bit side table (a folded bin): the reduce must
    # be 64-bit on every rank all the same
    total = reduce_dense_tables(lo, promote, int(t.max()), has_hi=(big == 2 and rank == 1))
    if total.dtype == torch.int32:
        res = total.numpy().view(np.uint32).astype(np.uint64)
    else:
        res = total.numpy().astype(np.uint64)
    np.save(os.path.join(tmp, f"res{rank}_{int(big)}.npy"), res)
    np.save(os.path.join(tmp, f"dtype{rank}_{int(big)}.npy"), np.array([total.element_size()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("big", [0, 1, 2])
def test_two_rank_shard_and_reduce(tmp_path, big):
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    from test_oracle_fastdense import SYNTH, dense_table, fastdense
    port = 29500 + (os.getpid() % 2000) + big
    mp.spawn(_worker, args=(2, port, str(tmp_path), big), nprocs=2, join=True)
    whole = subprocess.run([SYNTH, "0", str(N_REC), "3"], stdout=subprocess.PIPE, check=True).stdout
    want = dense_table(fastdense(), whole, K).astype(np.uint64)
    if big == 1:
        # what the per-rank tables summed to once bin 5 was overwritten on both ranks
        from kpop_b200.distributed import shard_records
        want[5] = 2 * 0xF0000000
    for r in range(2):
        got = np.load(tmp_path / f"res{r}_{int(big)}.npy")
        width = int(np.load(tmp_path / f"dtype{r}_{int(big)}.npy")[0])
        assert width == (8 if big else 4)  # a side table on ONE rank (big == 2) widens the reduce on all of them
        assert np.array_equal(got, want)


def test_shard_records_tiles_the_range():
    from kpop_b200.distributed import shard_records
    for total in (0, 1, 7, 8, 31_781_305):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                a, b = shard_records(total, world, r)
                assert a == prev and b >= a
                prev = b
            assert prev == total


def _range_worker(rank, world, port, path, tmp):
    sys.path.insert(0, ROOT)
    from kpop_b200.distributed import shard_fastq_byte_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_fastq_byte_range(path)
    np.save(os.path.join(tmp, f"range{rank}.npy"), np.array([a, b], dtype=np.int64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_fastq_file_is_cut_at_record_boundaries(tmp_path, world):
    """Read-chunk sharding of one FASTQ file: the per-rank byte ranges tile the file, every range starts
    on a line whose index is a multiple of 4 -- also when quality lines start with '@' or '+' -- and the per-range dense
    tables (CPU checker) add up to the table of the whole file."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    from test_oracle_fastdense import SYNTH, dense_table, fastdense
    data = subprocess.run([SYNTH, "0", "700", "3"], stdout=subprocess.PIPE, check=True).stdout
    trap = b"@x\nACGTACGTACGTACGT\n+\n@+@+@+@+@+@+@+@+\n@y\nTTGCACGTACGTAAAA\n+\n+@+@+@+@+@+@+@+@\n"
    data = trap * 40 + data + trap * 300 + b"@cut\nACGTACGTACGTAC"      # the last record is incomplete
    path = tmp_path / "reads.fq"
    path.write_bytes(data)
    port = 31000 + (os.getpid() % 2000) + world
    mp.spawn(_range_worker, args=(world, port, str(path), str(tmp_path)), nprocs=world, join=True)
    ranges = [tuple(int(x) for x in np.load(tmp_path / f"range{r}.npy")) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == len(data)
    for r in range(world - 1):
        assert ranges[r][1] == ranges[r + 1][0]
    lib = fastdense()
    total = np.zeros(4 ** K, dtype=np.uint64)
    for a, b in ranges:
        assert b >= a
        if 0 < a < len(data):
            assert data[a - 1:a] == b"\n" and data[:a].count(b"\n") % 4 == 0   # a record starts here
        total += dense_table(lib, data[a:b], K).astype(np.uint64)
    assert np.array_equal(total, dense_table(lib, data, K).astype(np.uint64))
    assert sum(1 for a, b in ranges if b > a) == world   # nobody idles on this input


# ---- sparse merge (last bullet): one large-k sample on two ranks, emulated library on the CPU -----------------
def _sparse_worker(rank, world, port, path, tmp, k):
    sys.path.insert(0, ROOT)
    from kpop_b200 import _native
    from kpop_b200.distributed import count_fastq_sharded_sparse
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["KPC_EMUL_TILE"] = "64x16"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _native.load(os.path.join(ROOT, "tests", "emul", "_build", "libkpopcount_emul.so"))
    text = count_fastq_sharded_sparse(path, k=k, label="x", lib=lib, chunk_bytes=20000)
    if rank == 0:
        with open(os.path.join(tmp, "sparse.txt"), "wb") as f:
            f.write(text)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,k", [(2, 21), (3, 13)])
def test_sparse_merge_of_a_sharded_sample_equals_one_process(tmp_path, world, k):
    """Each rank counts its shard on the hash-table path with stream-wide insertion ranks; the entries are exchanged by
    bucket owner, merged (kpc_hash_import) and dumped per owner: the concatenation must be what ONE KPopCount prints,
    order included (repeated reads make the same k-mer first appear on different ranks)."""
    import random
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emul")], check=True)
    rng = random.Random(17 + world)
    reads = [bytes(rng.choices(b"ACGTN", weights=[30, 30, 30, 30, 1], k=rng.choice([40, 80, 120]))) for _ in range(120)]
    recs = []
    for i in range(1500):
        s = rng.choice(reads)                   # heavy repetition: the same k-mers on every shard
        recs.append(b"@r%d\n%s\n+\n%s\n" % (i, s, bytes(rng.choice(b"@+I") for _ in range(len(s)))))
    path = tmp_path / "reads.fq"
    path.write_bytes(b"".join(recs))
    port = 33000 + (os.getpid() % 2000) + world
    mp.spawn(_sparse_worker, args=(world, port, str(path), str(tmp_path), k), nprocs=world, join=True)
    want = subprocess.run([os.path.join(ORACLE_DIR, "_build", "kpopcount_oracle"), "-k", str(k), "-l", "x", "-s", str(path)],
                          stdout=subprocess.PIPE, check=True).stdout
    assert (tmp_path / "sparse.txt").read_bytes() == want


# ---- a PAIR of FASTQ files on several ranks (both files cut at the same record index) ---------------------
def _pair_worker(rank, world, port, p1, p2, tmp, k):
    sys.path.insert(0, ROOT)
    from kpop_b200 import _native
    from kpop_b200.distributed import count_fastq_pair_sharded, pair_aligned_ranges
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["KPC_EMUL_TILE"] = "64x16"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _native.load(os.path.join(ROOT, "tests", "emul", "_build", "libkpopcount_emul.so"))
    r1, r2, first, n = pair_aligned_ranges(p1, p2)
    with open(os.path.join(tmp, "ranges.%d" % rank), "w") as f:
        f.write("%d %d %d %d %d %d\n" % (r1 + r2 + (first, n)))
    try:
        text = count_fastq_pair_sharded(p1, p2, k=k, label="x", lib=lib, chunk_bytes=3000)
        if rank == 0:
            with open(os.path.join(tmp, "pair.txt"), "wb") as f:
                f.write(text)
    except Exception as e:  # noqa: BLE001 -- every rank must see the same failure
        with open(os.path.join(tmp, "error.%d" % rank), "w") as f:
            f.write(str(e))
    dist.barrier()
    dist.destroy_process_group()


def _mates(rng, n, tag, unterminated=False, bad_at=None):
    out = []
    for i in range(n):
        ln = rng.randint(15, 70)
        seq = bytes(rng.choices(b"ACGTN", weights=[20, 20, 20, 20, 1], k=ln))
        head = b"@r%d/%d" % (i, tag) if i != bad_at else b"r%d/%d" % (i, tag)
        out.append(head + b"\n" + seq + b"\n+\n" + bytes(rng.choice(b"@+I") for _ in range(ln)) + b"\n")
    data = b"".join(out)
    return data[:-1] if unterminated and data else data


@pytest.mark.parametrize("world,n1,n2,unterminated,bad", [(2, 120, 90, False, None), (3, 75, 75, True, None), (3, 2, 40, False, None),
                                                          (2, 100, 130, False, (1, 77)), (3, 90, 60, True, (0, 5))])
def test_paired_files_cut_at_the_same_pair_on_every_rank(tmp_path, world, n1, n2, unterminated, bad):
    """pair_aligned_ranges: the ranks' ranges of the two files hold the same pairs, tile exactly the part of each file that
    FASTQ.iter_pe reads (up to the end of the shorter file's last complete record), and the summed dense tables print what
    ONE KPopCount -p prints -- including the line number of the first malformed pair."""
    import random
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emul")], check=True)
    rng = random.Random(world * 1000 + n1 + n2)
    m1 = _mates(rng, n1, 1, unterminated, bad[1] if bad and bad[0] == 0 else None)
    m2 = _mates(rng, n2, 2, unterminated, bad[1] if bad and bad[0] == 1 else None)
    p1, p2 = tmp_path / "m1.fq", tmp_path / "m2.fq"
    p1.write_bytes(m1); p2.write_bytes(m2)
    port = 35000 + (os.getpid() % 2000) + world + n1
    mp.spawn(_pair_worker, args=(world, port, str(p1), str(p2), str(tmp_path), 9), nprocs=world, join=True)
    ranges = [tuple(int(x) for x in open(tmp_path / ("ranges.%d" % r)).read().split()) for r in range(world)]
    pairs = min(n1, n2)
    assert ranges[0][0] == 0 and ranges[0][2] == 0 and ranges[0][4] == 0
    for r in range(world):
        s1, e1, s2, e2, first, n = ranges[r]
        assert m1[s1:e1].count(b"\n") + (1 if unterminated and e1 == len(m1) and e1 > s1 else 0) == 4 * n
        assert m2[s2:e2].count(b"\n") + (1 if unterminated and e2 == len(m2) and e2 > s2 else 0) == 4 * n
        if r + 1 < world:
            assert (e1, e2, first + n) == (ranges[r + 1][0], ranges[r + 1][2], ranges[r + 1][4])
        else:
            assert first + n == pairs
    oracle = subprocess.run([os.path.join(ORACLE_DIR, "_build", "kpopcount_oracle"), "-k", "9", "-l", "x", "-p", str(p1), str(p2)],
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if bad is None or bad[1] >= pairs:
        assert oracle.returncode == 0
        assert (tmp_path / "pair.txt").read_bytes() == oracle.stdout
    else:
        assert oracle.returncode == 2
        want_line = re.search(rb"On line (\d+)", oracle.stderr).group(1).decode()
        for r in range(world):
            assert ("On line %s:" % want_line) in open(tmp_path / ("error.%d" % r)).read()


# ---- -L (one spectrum per read) on several ranks --------------------------------------------------------------------------
def _per_record_worker(rank, world, port, path, tmp, k, m):
    sys.path.insert(0, ROOT)
    from kpop_b200 import _native
    from kpop_b200.distributed import count_fastq_sharded_per_record
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["KPC_EMUL_TILE"] = "64x16"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _native.load(os.path.join(ROOT, "tests", "emul", "_build", "libkpopcount_emul.so"))
    try:
        text = count_fastq_sharded_per_record(path, k=k, lib=lib, max_results_size=m, chunk_bytes=5000)
        if rank == 0:
            with open(os.path.join(tmp, "L.txt"), "wb") as f:
                f.write(text)
    except Exception as e:  # noqa: BLE001
        with open(os.path.join(tmp, "error.%d" % rank), "wb") as f:
            f.write(str(e).encode() + b"\n" + getattr(e, "partial_text", b""))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,bad,m", [(2, 150, None, 16777216), (3, 101, None, 4096), (3, 120, 83, 16777216), (2, 60, None, 8)])
def test_per_record_spectra_of_a_sharded_file(tmp_path, world, n, bad, m):
    """-L shards by records: the ranks' texts, concatenated in rank order, are what ONE KPopCount -L prints; a
    malformed record stops the run after the spectra before it, with the reference's line number; a table that could grow
    inside the run (tiny -M) is refused on every rank."""
    import random
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emul")], check=True)
    rng = random.Random(world * 100 + n)
    path = tmp_path / "reads.fq"
    path.write_bytes(_mates(rng, n, 1, False, bad))
    port = 37000 + (os.getpid() % 2000) + world + n
    mp.spawn(_per_record_worker, args=(world, port, str(path), str(tmp_path), 7, m), nprocs=world, join=True)
    oracle = subprocess.run([os.path.join(ORACLE_DIR, "_build", "kpopcount_oracle"), "-k", "7", "-L", "-M", str(m), "-s", str(path)],
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if m == 8:
        for r in range(world):
            assert b"could make the table" in open(tmp_path / ("error.%d" % r), "rb").read()
    elif bad is None:
        assert oracle.returncode == 0
        assert (tmp_path / "L.txt").read_bytes() == oracle.stdout
    else:
        assert oracle.returncode == 2
        want_line = re.search(rb"On line (\d+)", oracle.stderr).group(1)
        for r in range(world):
            msg, _, partial = open(tmp_path / ("error.%d" % r), "rb").read().partition(b"\n")
            assert b"On line " + want_line + b":" in msg
            assert partial == oracle.stdout          # the spectra the reference had printed before it failed
