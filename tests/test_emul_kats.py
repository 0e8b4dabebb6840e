"""CPU-side checks of the product's host engine + tile machine, run through the emulation harness.

tests/emul builds the *same* sources as libkpopcount_gpu.so (kpc_engine.cpp, kpc_abi.cpp, kpopcount_main.cpp and
the phase functions of kpc_tile.cuh) against a host-memory runtime, with tiny tiles and tiny staging buffers so
that every kind of boundary (thread segment, scan group, tile, launch) falls inside small inputs.  The oracle is
the checker.  The CUDA build of the same code is checked on a B200 by the -m gpu tests.
"""
import os
import subprocess

import pytest

from conftest import ROOT, run_cli
from kats import K9, KATS
from test_oracle_kats import materialise

EMUL_DIR = os.path.join(ROOT, "tests", "emul")
EMUL_BIN = os.path.join(EMUL_DIR, "_build", "KPopCount_emul")

GEOMETRIES = [("8x4", "64"), ("1x16", "97"), ("64x2", "4096"), ("256x64", "1000000")]


@pytest.fixture(scope="session")
def emul_bin():
    subprocess.run(["make", "-s", "-C", EMUL_DIR], check=True)
    return EMUL_BIN


def emul_env(tile, chunk):
    env = dict(os.environ)
    env["KPC_EMUL_TILE"] = tile
    env["KPC_CHUNK_BYTES"] = chunk
    return env


@pytest.mark.parametrize("geom", GEOMETRIES, ids=[g[0] + "_" + g[1] for g in GEOMETRIES])
@pytest.mark.parametrize("kat", KATS, ids=[k[0] for k in KATS])
def test_emul_kat(emul_bin, tmp_path, kat, geom):
    name, files, argv, expected, code = kat
    rc, out, err = run_cli(emul_bin, materialise(tmp_path, files, argv), env=emul_env(*geom))
    assert rc == code, err.decode(errors="replace")
    assert out == expected


@pytest.mark.parametrize("geom", GEOMETRIES, ids=[g[0] + "_" + g[1] for g in GEOMETRIES])
def test_emul_k9_bucket_order(emul_bin, tmp_path, geom):
    name, files, argv, head, code = K9
    rc, out, err = run_cli(emul_bin, materialise(tmp_path, files, argv), env=emul_env(*geom))
    assert rc == code, err.decode(errors="replace")
    assert out.split(b"\n")[: len(head)] == head


@pytest.mark.parametrize("label", ["", "all"], ids=["per_record", "one_label"])
def test_emul_sink_buffer_matches_callback(emul_bin, label):
    """kpc_set_sink_buffer (include/kpopcount.h): the text written into a caller-owned host buffer equals what the
    callback sink receives, in -L mode (records dumped from inside feed/end) and in -l mode (one dump at finish)."""
    import ctypes
    import subprocess
    import sys
    code = r'''
import ctypes, sys
sys.path.insert(0, %r)
from kpop_b200 import _native
lib = _native.load(%r)
from kpop_b200 import KMerCounter
data = b">a x\nACGTTGCANNACGATCGATCGGCTAGCTAGGATCG\nACGT\n>b\nTTTTTTTTGGGGGGGCCCCCCAAAAAAA\n>\nACGTACGTACGT\n>c\nacgtnacgtacgtagctagc\n" * 7
with KMerCounter(k=5, label=%r, lib=lib) as kc:
    kc.begin("fasta"); kc.feed(data, eof=True); kc.end(); kc.finish()
    want = kc.take_text()
    buf = ctypes.create_string_buffer(len(want) + 8)
    kc.reset()
    kc.set_text_buffer(ctypes.addressof(buf), len(buf))
    kc.begin("fasta"); kc.feed(data, eof=True); kc.end(); kc.finish()
    n = kc.text_buffer_used()
    assert n == len(want) and buf.raw[:n] == want and len(want) > 100, (n, len(want))
print("ok")
''' % (ROOT, os.path.join(EMUL_DIR, "_build", "libkpopcount_emul.so"), label)
    p = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0 and p.stdout.strip() == b"ok", p.stderr.decode(errors="replace")


def test_emul_cli_output_prefix_and_pipes(emul_bin, oracle_bin, tmp_path):
    """-o <prefix> / -o /dev/stdout (lib/KMerDB.ml:26-31) and `cat x.fa | KPopCount ... -f /dev/stdin` (README.md:89-96) through
    the same kpopcount_main.cpp the GPU binary is built from."""
    fa = tmp_path / "a.fa"
    fa.write_bytes(b">s one\nACGTNACGTTGACCA\nGGTAC\n>t\nacgtacgtacgt\n")
    env = emul_env("8x4", "64")
    rc, out, err = run_cli(emul_bin, ["-k", "3", "-l", "x", "-f", str(fa), "-o", str(tmp_path / "pre")], env=env)
    rc_o, out_o, _ = run_cli(oracle_bin, ["-k", "3", "-l", "x", "-f", str(fa)])
    assert rc == 0 and out == b"", err
    assert (tmp_path / "pre.KPopSpectra.txt").read_bytes() == out_o
    rc, out, err = run_cli(emul_bin, ["-k", "3", "-l", "x", "-f", str(fa), "-o", "/dev/stdout"], env=env)
    assert (rc, out) == (0, out_o), err
    for argv in (["-k", "5", "-L"], ["-k", "4", "-l", "w"], ["-k", "13", "-l", "w"]):
        rc_o, out_o, _ = run_cli(oracle_bin, argv + ["-f", str(fa)])
        rc_e, out_e, err_e = run_cli(emul_bin, argv + ["-f", "/dev/stdin"], stdin=fa.read_bytes(), env=env)
        assert (rc_e, out_e) == (rc_o, out_o), err_e


def _unequal_mates(seed, n1, n2):
    import random
    rng = random.Random(seed)
    def mate(n, tag, lo, hi):
        out = []
        for i in range(n):
            ln = rng.randint(lo, hi)
            out.append(b"@r%d/%d\n%s\n+\n%s\n" % (i, tag, bytes(rng.choices(b"ACGTN", weights=[9, 9, 9, 9, 1], k=ln)), b"I" * ln))
        return b"".join(out)
    return mate(n1, 1, 20, 60), mate(n2, 2, 10, 50)


@pytest.mark.parametrize("n1,n2", [(40, 25), (25, 40), (30, 30), (0, 5)])
def test_emul_cli_paired_fifos_one_pass(emul_bin, oracle_bin, tmp_path, n1, n2):
    """Paired-end input from two FIFOs with UNEQUAL numbers of records: FASTQ.iter_pe stops at the shorter file
    (Files.ml:228-247) and a pipe cannot be read again, so the CLI asks the library for one pass (kpc_set_single_pass);
    dense table (k = 5), hash table (k = 13) and -L."""
    m1, m2 = _unequal_mates(n1 * 100 + n2, n1, n2)
    f1, f2 = tmp_path / "m1.fq", tmp_path / "m2.fq"
    f1.write_bytes(m1); f2.write_bytes(m2)
    p1, p2 = str(tmp_path / "p1"), str(tmp_path / "p2")
    os.mkfifo(p1); os.mkfifo(p2)
    env = emul_env("8x4", "97")
    for argv in (["-k", "5", "-l", "x"], ["-k", "13", "-l", "x"], ["-k", "4", "-L"]):
        rc_o, out_o, _ = run_cli(oracle_bin, argv + ["-p", str(f1), str(f2)])
        # regular files: the two-pass route
        rc_f, out_f, err_f = run_cli(emul_bin, argv + ["-p", str(f1), str(f2)], env=env)
        assert (rc_f, out_f) == (rc_o, out_o), err_f
        writers = [subprocess.Popen(["sh", "-c", f"cat {f1} > {p1}"]), subprocess.Popen(["sh", "-c", f"cat {f2} > {p2}"])]
        rc_e, out_e, err_e = run_cli(emul_bin, argv + ["-p", p1, p2], env=env)
        for w in writers:
            w.wait()
        assert (rc_e, out_e) == (rc_o, out_o), err_e
