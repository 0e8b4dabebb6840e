"""The sort path for large-k samples (kpc_bucketsort.cuh + the bucket sinks of kpc_tile.cuh) under emulation, against the
oracle: genomes with repeated segments and low-complexity runs (groups for the CTA-wide kernel, groups too heavy for the
path -> fall back to the hash table), FASTQ, a second input after the first (migration into the hash table), small -M
(the path must step aside: the table can spill).  The emulation build uses tiny limits (3 pairs per thread, 12 per CTA)."""
import os
import random

import pytest

from conftest import run_cli
from fuzzgen import fasta, fastq
from test_emul_kats import emul_bin  # noqa: F401


def repeats(rng):
    unit = bytes(rng.choices(b"ACGT", k=rng.choice([20, 40, 80])))
    body = bytearray()
    for _ in range(rng.choice([3, 10, 40])):
        body += unit
        body += bytes(rng.choices(b"ACGT", k=rng.randrange(0, 30)))
        if rng.random() < 0.3:
            body += b"A" * rng.choice([35, 80])
    out = bytearray(b">g\n")
    w = rng.choice([60, 70, 10 ** 6])
    for i in range(0, len(body), w):
        out += body[i:i + w] + b"\n"
    return bytes(out)


@pytest.mark.parametrize("seed", range(4))
def test_emul_sort_path_vs_oracle(emul_bin, oracle_bin, tmp_path, seed):
    for c in range(int(os.environ.get("KPC_FUZZ_CASES", "30"))):
        rng = random.Random(90000 + 1000 * seed + c)
        k = rng.choice([13, 14, 16, 21, 30])
        content = rng.choice(["DNA-ds", "DNA-ds", "DNA-ss"])
        kind = rng.choice(["fasta", "repeats", "repeats", "fastq", "two"])
        argv = ["-k", str(k), "-C", content, "-l", "x"]
        f1, f2 = tmp_path / f"c{c}_1", tmp_path / f"c{c}_2"
        if kind == "fasta":
            f1.write_bytes(fasta(rng, False, max_records=6, max_len=400)); argv += ["-f", str(f1)]
        elif kind == "repeats":
            f1.write_bytes(repeats(rng)); argv += ["-f", str(f1)]
        elif kind == "fastq":
            f1.write_bytes(fastq(rng, False, max_records=20, malformed=rng.choice([0, 0, 0.05]))); argv += ["-s", str(f1)]
        else:
            f1.write_bytes(repeats(rng)); f2.write_bytes(fasta(rng, False, max_records=4, max_len=300))
            argv += ["-f", str(f1), "-f", str(f2)]
        if rng.random() < 0.3:
            argv += ["-M", str(rng.choice([50, 1000, 100000]))]
        env = dict(os.environ)
        env["KPC_EMUL_TILE"] = rng.choice(["64x16", "8x4", "256x64"])
        env["KPC_CHUNK_BYTES"] = str(rng.choice([1 << 20, 1 << 20, 4096]))
        rc_o, out_o, _ = run_cli(oracle_bin, argv)
        rc_e, out_e, err_e = run_cli(emul_bin, argv, env=env)
        ctx = f"seed={seed} case={c} kind={kind} argv={' '.join(argv)} env={env['KPC_EMUL_TILE']},{env['KPC_CHUNK_BYTES']}\n" \
              f"{err_e.decode(errors='replace')[-300:]}"
        assert "code -9" not in err_e.decode(errors="replace"), ctx
        assert (rc_e, out_e) == (rc_o, out_o), ctx


def test_sort_path_is_taken_and_steps_aside(emul_bin, oracle_bin, tmp_path):
    """Kernel-launch counts tell the paths apart: a plain genome stays on the sort path, one with a long homopolymer
    run falls back to the hash table (same text either way)."""
    rng = random.Random(3)
    plain = tmp_path / "plain.fa"
    plain.write_bytes(b">g\n" + bytes(rng.choices(b"ACGT", k=900)) + b"\n")
    heavy = tmp_path / "heavy.fa"
    heavy.write_bytes(b">g\n" + bytes(rng.choices(b"ACGT", k=300)) + b"A" * 200 + b"\n")

    def launches(path, sort):
        env = dict(os.environ, KPC_SORT_PATH=sort, KPC_EMUL_TILE="64x16")
        rc, out, err = run_cli(emul_bin, ["-k", "21", "-l", "x", "-v", "-f", str(path)], env=env)
        assert rc == 0
        rc_o, out_o, _ = run_cli(oracle_bin, ["-k", "21", "-l", "x", "-f", str(path)])
        assert out == out_o
        return int(err.decode().strip().split("\n")[-1].split()[1])

    assert launches(plain, "1") != launches(plain, "0")
    assert launches(heavy, "1") > launches(plain, "1")  # sort path tried first, then the hash table


def test_batch_of_samples_through_one_context(emul_bin, oracle_bin, tmp_path):
    """kpc_reset_label + kpc_feed_device on the hash-table path: many samples, one context (the C4 idiom), each spectrum equal
    to what a separate `KPopCount -l <label>` run prints."""
    from test_emul_fastpath import DRIVER
    rng = random.Random(11)
    files, want = [], b""
    for i in range(5):
        p = tmp_path / f"g{i}.fa"
        p.write_bytes(repeats(rng) if i % 2 else b">g\n" + bytes(rng.choices(b"ACGT", k=700)) + (b"\n" if i != 2 else b""))
        files.append(str(p))
        rc, out, _ = run_cli(oracle_bin, ["-k", "21", "-l", f"S{i}", "-f", str(p)])
        assert rc == 0
        want += out
    env = dict(os.environ, KPC_DRIVER_BATCH="1", KPC_EMUL_TILE="64x16")
    rc, out, err = run_cli(DRIVER, ["21", "DNA-ds", "S", "fasta"] + files, env=env)
    assert rc == 0, err
    assert out == want
