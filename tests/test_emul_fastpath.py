"""The dense-table pipeline (kpc_partition.cuh: partition + count kernels) under the SIMT emulator, against the oracle.

tests/emul compiles the very kernel source nvcc builds for sm_100a with g++ (KPC_SIMT_EMUL) and runs it with cooperative
fibers: warps, barriers, shuffles, shared memory and the decoupled look-back across CTAs behave as on the device, with
tiny geometries (KPC_EMUL_FQ=threads x vectors-per-thread x CTAs) so that tile, batch, bucket and launch boundaries fall
everywhere in small inputs.  Inputs go through the CLI (kpc_feed, launches cut at line starts) and through
kpc_feed_device (consecutive launches of one buffer that read the 16 bytes before them).
"""
import os
import random
import subprocess

import pytest

from conftest import run_cli
from fuzzgen import fastq
from test_emul_kats import EMUL_DIR, emul_bin  # noqa: F401

DRIVER = os.path.join(EMUL_DIR, "_build", "device_feed_driver")
UNIT = os.path.join(EMUL_DIR, "_build", "unit_checks")
GEOMS = ["32x1", "32x1x1", "32x2", "64x1", "64x4", "128x4", "32x1x7"]


def regular(rng, n, length, ragged=False, crlf=False):
    out = bytearray()
    for i in range(n):
        ln = rng.randrange(max(1, length // 2), length + 1) if ragged else length
        seq = bytes(rng.choices(b"ACGTNacgt", weights=[25, 25, 25, 25, 1, 2, 2, 2, 2], k=ln))
        eol = b"\r\n" if crlf else b"\n"
        out += b"@r%d" % i + eol + seq + eol + b"+" + eol + bytes(rng.choice(b"@+I5#") for _ in range(ln)) + eol
    return bytes(out)


def test_simd_helpers_exhaustive(emul_bin):
    """fq_classify4 on every byte value in every lane, fq_nl_mask16 on random vectors: against byte-at-a-time code."""
    p = subprocess.run([UNIT], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()


@pytest.mark.parametrize("seed", range(8))
def test_emul_fast_path_vs_oracle(emul_bin, oracle_bin, tmp_path, seed):
    n_cases = int(os.environ.get("KPC_FUZZ_CASES", "30"))
    skipped = 0
    for c in range(n_cases):
        rng = random.Random(7000 + 1000 * seed + c)
        k = rng.choice([4, 5, 6, 7, 8, 9, 10, 11, 12, 12, 12])
        content = rng.choice(["DNA-ds", "DNA-ds", "DNA-ss"])
        kind = rng.choice(["fuzz", "fuzz", "regular", "ragged", "long", "skew"])

        def gen():
            if kind == "fuzz":
                return fastq(rng, False, max_records=rng.choice([3, 10, 40]), malformed=rng.choice([0, 0, 0, 0.05]))
            if kind == "regular":
                return regular(rng, rng.choice([1, 5, 30, 120]), rng.choice([20, 36, 100, 150]), crlf=rng.random() < 0.1)
            if kind == "ragged":
                return regular(rng, rng.choice([5, 30, 100]), rng.choice([30, 150, 400]), ragged=True)
            if kind == "skew":  # the same few k-mers over and over: buckets and queues overflow
                out = bytearray()
                for i in range(rng.choice([20, 100, 300])):
                    unit = rng.choice([b"A", b"A", b"AC", b"ACG", b"T", b"GGGC"])
                    ln = rng.choice([50, 150, 151])
                    out += b"@s%d\n" % i + (unit * (ln // len(unit) + 1))[:ln] + b"\n+\n" + b"I" * ln + b"\n"
                return bytes(out)
            return regular(rng, rng.choice([1, 3]), rng.choice([1000, 5000, 20000]))

        pe = rng.random() < 0.25
        argv = ["-k", str(k), "-C", content, "-l", "x"]
        f1, f2 = str(tmp_path / f"c{c}_1.fq"), str(tmp_path / f"c{c}_2.fq")
        with open(f1, "wb") as f:
            f.write(gen())
        if pe:
            with open(f2, "wb") as f:
                f.write(gen())
            argv += ["-p", f1, f2]
        else:
            argv += ["-s", f1]
        env = dict(os.environ)
        env["KPC_EMUL_FQ"] = rng.choice(GEOMS)
        env["KPC_EMUL_TILE"] = "64x16"
        env["KPC_CHUNK_BYTES"] = str(rng.choice([4096, 5000, 8192, 70000, 1 << 20]))
        env["KPC_FQ_LAUNCH_BYTES"] = str(rng.choice([512, 2048, 1 << 20]))
        devices = rng.choice([None, None, "0,1", "0,1,2"])  # several (emulated) devices behind one context
        if devices:
            env["KPC_DEVICES"] = devices
        if rng.random() < 0.3:  # queues that overflow on small inputs: the count-in-place paths
            env["KPC_FQ_QUEUE_PERMILLE"] = str(rng.choice([0, 100, 400]))
            env["KPC_FQ_QUEUE_SLACK"] = str(rng.choice([0, 16, 64]))
        rc_o, out_o, _ = run_cli(oracle_bin, argv)
        if not pe and rng.random() < 0.5:
            rc_e, out_e, err_e = run_cli(DRIVER, [str(k), content, "x", "single-end", f1], env=env)
        else:
            rc_e, out_e, err_e = run_cli(emul_bin, argv, env=env)
        ctx = f"seed={seed} case={c} kind={kind} env={env['KPC_EMUL_FQ']},{env['KPC_CHUNK_BYTES']},{env['KPC_FQ_LAUNCH_BYTES']},{env.get('KPC_DEVICES')} " \
              f"argv={' '.join(argv)}\n{err_e.decode(errors='replace')}"
        if rc_e == 2 and b"code -9" in err_e:
            skipped += 1  # KPC_E_UNSUPPORTED: lines longer than the staging size (documented refusal)
            continue
        assert rc_e == rc_o, ctx
        assert out_e == out_o, ctx
    assert skipped <= n_cases // 4, f"{skipped} of {n_cases} cases were refused"
