"""The fast multi-threaded CPU counter (oracle/fast_dense.cpp, the checker for large parity runs) agrees with the
faithful oracle, and the host generator of the synthetic reads has the frozen shape."""
import ctypes
import hashlib
import os
import random
import subprocess

import numpy as np

from conftest import ORACLE_DIR, run_cli
from fuzzgen import fastq

LIB = os.path.join(ORACLE_DIR, "_build", "libkpc_fastdense.so")
SYNTH = os.path.join(ORACLE_DIR, "_build", "synth_fastq")


def fastdense():
    lib = ctypes.CDLL(LIB)
    lib.fd_count_fastq_dense.restype = ctypes.c_uint64
    lib.fd_count_fastq_dense.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    lib.fd_synth_valid_windows.restype = ctypes.c_uint64
    lib.fd_synth_valid_windows.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int]
    return lib


def dense_table(lib, data, k, threads=4):
    t = np.zeros(4 ** k, dtype=np.uint32)
    lib.fd_count_fastq_dense(data, len(data), k, t.ctypes.data_as(ctypes.c_void_p), threads)
    return t


def table_to_text(t, k, label):
    w = (2 * k + 3) // 4
    nz = np.nonzero(t)[0]
    return ("\t%s\n" % label + "".join("%0*x\t%d\n" % (w, i, t[i]) for i in nz)).encode()


def test_fastdense_matches_oracle_on_fuzzed_fastq(oracle_bin, tmp_path):
    lib = fastdense()
    rng = random.Random(7)
    for i in range(60):
        data = fastq(rng, max_records=12)
        k = rng.choice([1, 2, 3, 5, 8, 11])
        p = tmp_path / f"f{i}.fq"
        p.write_bytes(data)
        rc, out, err = run_cli(oracle_bin, ["-k", str(k), "-l", "x", "-s", str(p)])
        if rc != 0:
            continue  # malformed inputs are the faithful oracle's business only
        assert table_to_text(dense_table(lib, data, k), k, "x") == out


def test_synth_generator_shape_and_window_count(oracle_bin, tmp_path):
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    data = subprocess.run([SYNTH, "0", "3000", "3"], stdout=subprocess.PIPE, check=True).stdout
    # frozen stream: any change of kpc_synth.h changes every benchmark number
    assert len(data) == 307 * 3000 + (10 * 1 + 90 * 2 + 900 * 3 + 2000 * 4)
    assert hashlib.md5(data).hexdigest() == "9f85d52bb410e40169066884360eaddd"
    assert data.startswith(b"@S0\n") and data.count(b"\n") == 12000
    lib = fastdense()
    t = dense_table(lib, data, 12)
    assert int(t.sum()) == lib.fd_synth_valid_windows(0, 3000, 3, 12, 3)
    p = tmp_path / "s.fq"
    p.write_bytes(data)
    rc, out, _ = run_cli(oracle_bin, ["-k", "12", "-l", "S", "-s", str(p)])
    assert rc == 0 and table_to_text(t, 12, "S") == out
    # a later slice of the stream is the same bytes
    part = subprocess.run([SYNTH, "1000", "500", "3"], stdout=subprocess.PIPE, check=True).stdout
    off = 307 * 1000 + 10 + 180 + 2700
    assert data[off:off + len(part)] == part
