"""Parity of the fast FASTQ pipeline (kpop_b200/csrc/kpc_fastq.cu: partition + shared-memory counting) against the
oracle, bit-exact, through the C ABI (Python binding), on adversarial FASTQ that crosses many tiles, launches and
line batches.  Every case is also run with KPC_FAST=0 (the generic tile machine) so that a disagreement can be
attributed.  Run on a B200 with `pytest -m gpu`."""
import os
import random
import zlib

import pytest

from conftest import run_cli
from fuzzgen import fastq, rand_name, rand_seq

pytestmark = pytest.mark.gpu


def big_fastq(rng, shape):
    """A few hundred kB of FASTQ of a given shape (tiles are 32 KiB, line batches 2048 lines)."""
    out = bytearray()
    if shape == "reads150":
        for i in range(rng.randrange(1500, 2500)):
            n = rng.choice([150, 150, 150, 151, 100, 36])
            seq = bytes(rng.choices(b"ACGTNacgt", weights=[30, 30, 30, 30, 1, 2, 2, 2, 2], k=n))
            out += b"@r%d\n%s\n+\n%s\n" % (i, seq, bytes(rng.choice(b"@+I5#") for _ in range(n)))
    elif shape == "tiny_lines":  # thousands of lines per tile: several line batches
        for i in range(rng.randrange(8000, 16000)):
            n = rng.choice([0, 0, 1, 2, 3, 5, 13, 20])
            out += b"@" + rand_name(rng).replace(b"\n", b"")[:2] + b"\n" + rand_seq(rng, n, False).replace(b"\n", b"") + b"\n+\n" + b"I" * rng.choice([0, n]) + b"\n"
    elif shape == "long_reads":  # lines longer than a tile
        for i in range(rng.randrange(3, 8)):
            n = rng.choice([10, 5000, 33000, 70000, 140000])
            seq = bytes(rng.choices(b"ACGTN", weights=[30, 30, 30, 30, 1], k=n))
            out += b"@long%d\n%s\n+\n%s\n" % (i, seq, b"I" * rng.choice([n, 7]))
    elif shape == "crlf":
        for i in range(rng.randrange(1000, 2000)):
            n = rng.randrange(0, 200)
            seq = bytes(rng.choices(b"ACGT", k=n))
            out += b"@r%d\r\n%s\r\n+\r\n%s\r\n" % (i, seq, b"I" * n)
    elif shape == "skewed":  # one slice takes almost everything: the queue overflows into the in-place path
        for i in range(rng.randrange(1500, 2500)):
            n = rng.choice([150, 250])
            seq = bytes(rng.choices(b"AT", weights=[50, 1], k=n)) if rng.random() < 0.9 else bytes(rng.choices(b"ACGT", k=n))
            out += b"@p%d\n%s\n+\n%s\n" % (i, seq, b"I" * n)
    else:  # "adversarial": the small-case generator, many records
        out += fastq(rng, False, max_records=3000, max_len=300, malformed=0.0)
    r = rng.random()
    if out and r < 0.2:
        out = out[:-1]
    elif out and r < 0.4:
        out = out[: len(out) - rng.choice([1, 2, 5, 30, 160, 400])]
    return bytes(out)


def count_with_binding(data, k, content, label, chunk, fast, device_feed=False, launch_bytes=None):
    import torch
    from kpop_b200 import KMerCounter
    from kpop_b200.counter import Content, KPopCountError
    old = {v: os.environ.get(v) for v in ("KPC_FAST", "KPC_CHUNK_BYTES", "KPC_FQ_LAUNCH_BYTES")}
    os.environ["KPC_FAST"] = "1" if fast else "0"
    if chunk:
        os.environ["KPC_CHUNK_BYTES"] = str(chunk)
    if launch_bytes:
        os.environ["KPC_FQ_LAUNCH_BYTES"] = str(launch_bytes)
    try:
        with KMerCounter(k=k, content=Content.of_string(content), label=label) as kc:
            try:
                kc.begin("single-end")
                if device_feed:
                    dev = torch.zeros(len(data) + 64, dtype=torch.uint8, device="cuda")
                    if data:
                        dev[: len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
                    kc.feed_device(dev.data_ptr(), len(data), eof=True)
                else:
                    step = chunk or (1 << 20)
                    pos = 0
                    while True:
                        piece = data[pos: pos + step]
                        pos += step
                        eof = pos >= len(data)
                        kc.feed(piece, eof=eof)
                        if eof:
                            break
                kc.end()
                kc.finish()
                return 0, kc.take_text()
            except KPopCountError as e:
                return e.code, kc.take_text()
    finally:
        for v, x in old.items():
            if x is None:
                os.environ.pop(v, None)
            else:
                os.environ[v] = x


SHAPES = ["reads150", "tiny_lines", "long_reads", "crlf", "skewed", "adversarial"]


@pytest.mark.parametrize("shape", SHAPES)
def test_fast_fastq_vs_oracle(oracle_bin, tmp_path, shape):
    rng = random.Random(zlib.crc32(shape.encode()))
    for rep in range(3):
        data = big_fastq(rng, shape)
        p = tmp_path / f"{shape}{rep}.fq"
        p.write_bytes(data)
        k = rng.choice([12, 12, 11, 9, 8, 5, 4])
        content = rng.choice(["DNA-ds", "DNA-ds", "DNA-ss"])
        if content == "DNA-ss" and k == 12:
            k = 11  # DNA-ss k = 12 can spill: hash path, not this pipeline
        rc_o, out_o, _ = run_cli(oracle_bin, ["-k", str(k), "-C", content, "-l", "x", "-s", str(p)])
        assert rc_o == 0
        for chunk, device_feed, launch in [(None, False, None), (65536, False, None), (None, True, 65536), (None, True, None)]:
            rc_f, out_f = count_with_binding(data, k, content, "x", chunk, True, device_feed, launch)
            ctx = f"shape={shape} rep={rep} k={k} {content} chunk={chunk} device_feed={device_feed} launch={launch} bytes={len(data)}"
            if rc_f == -9:
                continue  # KPC_E_UNSUPPORTED: an explicit refusal (lines longer than the staging size), never a wrong answer
            if (rc_f, out_f) != (0, out_o):
                rc_g, out_g = count_with_binding(data, k, content, "x", chunk, False, device_feed, launch)
                assert False, ctx + f": fast path differs from the oracle (generic path {'agrees' if out_g == out_o else 'differs too'})"


def test_fast_fastq_malformed_and_traps(oracle_bin, tmp_path):
    """'@' / '+' checks of FASTQ.iter_se (Files.ml:213-214) on the fast path: first bad record wins, a truncated last
    record is never checked, quality lines starting with '@' are not tags."""
    rng = random.Random(77)
    recs = [b"@r%d\nACGTACGTACGTACGTAC\n+\n@@@@++++IIII@@@@++\n" % i for i in range(4000)]
    good = b"".join(recs)
    cases = [good, b"".join(recs[:1500]) + b"xr\nACGT\n+\nIIII\n" + b"".join(recs[1500:]),
             good + b"@last\nACGTACGTACGTACGT\n-\nIIII\n", good + b"@last\nACGTACGTACGTACGT\n-\nII",
             good + b"\n", good.replace(b"@r2000\n", b"\n", 1), good.replace(b"+\n@@@@", b"\n@@@@", 1)]
    for i, data in enumerate(cases):
        p = tmp_path / f"m{i}.fq"
        p.write_bytes(data)
        rc_o, out_o, _ = run_cli(oracle_bin, ["-k", "12", "-l", "x", "-s", str(p)])
        rc_f, out_f = count_with_binding(data, 12, "DNA-ds", "x", rng.choice([None, 65536]), True)
        if rc_o == 0:
            assert (rc_f, out_f) == (0, out_o), i
        else:
            assert rc_o == 2 and rc_f == -3, (i, rc_o, rc_f)  # KPC_E_MALFORMED_FASTQ; stdout holds the header only
            assert out_f == out_o == b"\tx\n", i


def test_sink_buffer_matches_callback():
    """kpc_set_sink_buffer (text written straight into a caller-owned host buffer) gives the bytes of the callback sink."""
    import ctypes
    from kpop_b200 import KMerCounter
    rng = random.Random(77)
    data = big_fastq(rng, "reads150")
    with KMerCounter(k=12, label="buf") as kc:
        kc.begin("single-end"); kc.feed(data, eof=True); kc.end(); kc.finish()
        want = kc.take_text()
        buf = ctypes.create_string_buffer(len(want) + 64)
        kc.reset()
        kc.set_text_buffer(ctypes.addressof(buf), len(buf))
        kc.begin("single-end"); kc.feed(data, eof=True); kc.end(); kc.finish()
        n = kc.text_buffer_used()
        assert n == len(want) and buf.raw[:n] == want
        kc.reset()
        kc.set_text_buffer(ctypes.addressof(buf), 16)   # too small: a clean error, not an overrun
        kc.begin("single-end"); kc.feed(data, eof=True); kc.end()
        with pytest.raises(Exception):
            kc.finish()


def test_profile_hooks_report_the_fast_kernels():
    """kpc_profile_enable / kpc_profile_read (measurement aid of bench.py): CUDA-event times of the partition and count
    kernels, summed over the launches since the last read."""
    import torch
    from kpop_b200 import KMerCounter
    rng = random.Random(5)
    data = big_fastq(rng, "reads150")
    with KMerCounter(k=12, label="p") as kc:
        kc.profile_enable(True)
        dev = torch.zeros(len(data) + 64, dtype=torch.uint8, device="cuda")
        dev[: len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
        kc.begin("single-end"); kc.feed_device(dev.data_ptr(), len(data), eof=True); kc.end()
        part_ms, count_ms, launches, nbytes = kc.profile_read()
        assert launches >= 1 and part_ms > 0 and count_ms > 0 and 0 < nbytes <= len(data) + 1
        assert kc.profile_read()[2] == 0   # reading resets
        kc.profile_enable(False)
        kc.finish()


@pytest.mark.parametrize("k", [12, 7])
def test_fast_fastq_paired_end_big(oracle_bin, tmp_path, k):
    """-p with two large mate files of different record counts (the shorter one cut inside a record): FASTQ.iter_pe
    (Files.ml:222-250) stops at the last complete pair, which reaches the fast kernel as a line cap (max_lines)."""
    from test_gpu_parity import GPU_BIN
    rng = random.Random(4242 + k)

    def mate(n_reads, tag):
        out = bytearray()
        for i in range(n_reads):
            n = rng.choice([150, 150, 151, 100, 36, 250])
            seq = bytes(rng.choices(b"ACGTNacgt", weights=[30, 30, 30, 30, 1, 2, 2, 2, 2], k=n))
            out += b"@r%d/%s\n%s\n+\n%s\n" % (i, tag, seq, bytes(rng.choice(b"@+I5#ACGT") for _ in range(n)))
        return bytes(out)

    m1, m2 = mate(2600, b"1"), mate(2300, b"2")
    m2 = m2[: len(m2) - 77]   # the last record of mate 2 is incomplete: 2299 complete pairs
    f1, f2 = tmp_path / "m1.fq", tmp_path / "m2.fq"
    f1.write_bytes(m1); f2.write_bytes(m2)
    argv = ["-k", str(k), "-l", "pe", "-p", str(f1), str(f2)]
    rc_o, out_o, _ = run_cli(oracle_bin, argv)
    assert rc_o == 0
    for chunk in (None, "65536"):
        env = dict(os.environ)
        if chunk:
            env["KPC_CHUNK_BYTES"] = chunk
        rc_g, out_g, err_g = run_cli(GPU_BIN, argv, env=env)
        assert rc_g == 0, err_g.decode(errors="replace")
        assert out_g == out_o, f"paired-end k={k} chunk={chunk}"


def test_count_fastq_sharded_single_rank(oracle_bin, tmp_path):
    """distributed.count_fastq_sharded without a process group (one rank owns the whole file): the same bytes as the oracle;
    the multi-rank cut logic is covered on CPU by tests/test_distributed_gloo.py and on 2 GPUs by tools/mgpu_file_check.py."""
    from kpop_b200.distributed import count_fastq_sharded, shard_fastq_byte_range
    rng = random.Random(99)
    data = big_fastq(rng, "reads150") + b"@cut\nACGTACGTACGTAC"
    p = tmp_path / "one.fq"
    p.write_bytes(data)
    assert shard_fastq_byte_range(str(p)) == (0, len(data))
    rc_o, out_o, _ = run_cli(oracle_bin, ["-k", "12", "-l", "x", "-s", str(p)])
    assert rc_o == 0
    assert count_fastq_sharded(str(p), k=12, label="x", chunk_bytes=100_000) == out_o
