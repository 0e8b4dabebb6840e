"""kpc_set_single_pass through the Python binding on the emulated library: paired files of unequal length counted in ONE
pass (FASTQ.iter_pe stops at the shorter file, Files.ml:228-247), whatever order and chunking the caller feeds the two
mates in, with one or several (emulated) devices behind the context."""
import os
import random
import subprocess
import sys

import pytest

from conftest import ROOT, run_cli
from test_emul_kats import EMUL_DIR, _unequal_mates

_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
from kpop_b200 import _native
from kpop_b200.counter import KMerCounter
lib = _native.load(os.path.join(sys.argv[1], "tests", "emul", "_build", "libkpopcount_emul.so"))
k, order, chunk, ndev = int(sys.argv[4]), sys.argv[5], int(sys.argv[6]), int(sys.argv[7])
m = [open(sys.argv[2], "rb").read(), open(sys.argv[3], "rb").read()]
kc = KMerCounter(k=k, label="x", lib=lib, devices=list(range(ndev)))
kc.set_single_pass(True)
kc.begin("paired-end")
if order == "sequential":          # all of mate 1, then all of mate 2
    for i in (0, 1):
        pieces = [m[i][o:o + chunk] for o in range(0, len(m[i]), chunk)] or [b""]
        for j, p in enumerate(pieces):
            kc.feed(p, mate=i, eof=j == len(pieces) - 1)
else:                              # a chunk of each in turn, the longer file on its own at the end
    off, done = [0, 0], [False, False]
    while not all(done):
        for i in (0, 1):
            if done[i]:
                continue
            p = m[i][off[i]:off[i] + chunk]
            off[i] += len(p)
            done[i] = off[i] >= len(m[i])
            kc.feed(p, mate=i, eof=done[i])
kc.end()
kc.finish()
sys.stdout.buffer.write(kc.take_text())
"""


@pytest.mark.parametrize("k", [5, 13])
@pytest.mark.parametrize("order", ["sequential", "alternating"])
def test_single_pass_api_unequal_mates(oracle_bin, tmp_path, k, order):
    subprocess.run(["make", "-s", "-C", EMUL_DIR], check=True)
    rng = random.Random(k * 7 + len(order))
    n1, n2 = rng.choice([(40, 27), (19, 33)])
    m1, m2 = _unequal_mates(n1 * 100 + n2 + k, n1, n2)
    f1, f2 = tmp_path / "m1.fq", tmp_path / "m2.fq"
    f1.write_bytes(m1); f2.write_bytes(m2)
    rc_o, want, _ = run_cli(oracle_bin, ["-k", str(k), "-l", "x", "-p", str(f1), str(f2)])
    assert rc_o == 0
    for ndev in (1, 2):
        env = dict(os.environ, KPC_EMUL_TILE="8x4", KPC_CHUNK_BYTES="8192")
        p = subprocess.run([sys.executable, "-c", _WORKER, ROOT, str(f1), str(f2), str(k), order, str(rng.choice([97, 1000, 5000])),
                            str(ndev)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
        assert p.returncode == 0, p.stderr.decode(errors="replace")[-400:]
        assert p.stdout == want, (ndev, p.stderr.decode(errors="replace")[-200:])
