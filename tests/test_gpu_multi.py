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
@pytest.mark.parametrize("tool,arg", [("mgpu_file_check.py", "300000"), ("mgpu_sparse_check.py", "60000")],
                         ids=["dense_table_reduce", "sparse_merge"])
def test_one_fastq_file_on_two_gpus(tool, arg, oracle_bin):
    """Read-chunk sharding of one file on two ranks: dense tables + NCCL reduce (k = 12), and the sparse merge of the
    hash-table path (k = 21: entries exchanged by bucket owner, merged on the device)."""
    world = 2
    port = 29700 + os.getpid() % 200 + (7 if "sparse" in tool else 0)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tools", tool), arg],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    out = p.stdout.decode(errors="replace")
    assert p.returncode == 0, out[-2000:]
    assert "IDENTICAL TO ORACLE" in out, out[-2000:]


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_two_gpus_behind_one_context_of_the_c_abi(oracle_bin, tmp_path):
    """kpc_create(n_devices = 2): the chunks of a FASTQ stream go to the two devices in turn, kpc_finish adds the tables up
    over a peer copy.  Through the CLI (KPC_DEVICES=0,1; argv unchanged) and through the binding, against the oracle."""
    import random
    from conftest import run_cli
    from kpop_b200 import KMerCounter
    synth = os.path.join(ROOT, "oracle", "_build", "synth_fastq")
    fq = tmp_path / "reads.fq"
    trap = b"@x\nACGTACGTACGTACGT\n+\n@+@+@+@+@+@+@+@+\n@y\nTTGCACGTACGTAAAA\n+\n+@+@+@+@+@+@+@+@\n"
    with open(fq, "wb") as f:
        f.write(trap * 500)
        subprocess.run([synth, "0", "400000", "3"], stdout=f, check=True)
        f.write(trap * 500 + b"@cut\nACGTACGTACGTAC")
    gpu_bin = os.path.join(ROOT, "kpop_b200", "bin", "KPopCount")
    for k, content in ((12, "DNA-ds"), (9, "DNA-ss")):
        argv = ["-k", str(k), "-C", content, "-l", "x", "-s", str(fq)]
        rc_o, out_o, _ = run_cli(oracle_bin, argv)
        for chunk in ("8388608", "1000003"):
            env = dict(os.environ, KPC_DEVICES="0,1", KPC_CHUNK_BYTES=chunk)
            rc_g, out_g, err_g = run_cli(gpu_bin, argv, env=env)
            assert (rc_g, out_g) == (rc_o, out_o), err_g.decode(errors="replace")[-400:]
    # paired-end through the binding, both mates spread over the two devices; a hash-table run stays on the first device
    rng = random.Random(8)
    m = [b"".join(b"@p%d/%d\n%s\n+\n%s\n" % (i, j, bytes(rng.choices(b"ACGTN", weights=[30, 30, 30, 30, 1], k=120)), b"I" * 120)
                  for i in range(30000)) for j in (1, 2)]
    p1, p2 = tmp_path / "m1.fq", tmp_path / "m2.fq"
    p1.write_bytes(m[0]); p2.write_bytes(m[1])
    for k in (12, 15):
        rc_o, out_o, _ = run_cli(oracle_bin, ["-k", str(k), "-l", "pe", "-p", str(p1), str(p2)])
        with KMerCounter(k=k, label="pe", devices=[0, 1]) as kc:
            got = kc.compute([("paired-end", str(p1), str(p2))], chunk_bytes=1 << 20)
        assert got == out_o
