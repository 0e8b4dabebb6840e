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
