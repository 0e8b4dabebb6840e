"""Multi-GPU paths on a box with at least two GPUs (skipped otherwise): read-chunk sharding of ONE FASTQ file with the
line-feed census on the device and an NCCL table reduce (kpop_b200.distributed.count_fastq_sharded), byte-compared with
the oracle; launched the way bench.py is (torch.distributed.run, one rank per GPU)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world", [2])
def test_one_fastq_file_on_two_gpus(world, oracle_bin):
    port = 29700 + os.getpid() % 200
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tools", "mgpu_file_check.py"), "300000"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    out = p.stdout.decode(errors="replace")
    assert p.returncode == 0, out[-2000:]
    assert "IDENTICAL TO ORACLE" in out, out[-2000:]
