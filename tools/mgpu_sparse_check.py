#!/usr/bin/env python
"""Under torchrun: ONE large-k sample (k = 21, hash-table path) counted by all ranks with the sparse merge of
kpop_b200.distributed.count_fastq_sharded_sparse and compared byte for byte with the oracle on rank 0.
usage: torchrun ... tools/mgpu_sparse_check.py [reads]"""
import os, random, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from kpop_b200.distributed import count_fastq_sharded_sparse

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000
path = "/dev/shm/kpc_mgpu_sparse.fq"
if rank == 0:
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    rng = random.Random(5)
    genome = bytes(rng.choices(b"ACGT", k=200_000))       # reads of a small genome: every k-mer shows up on every rank
    with open(path, "wb") as f:
        for i in range(reads):
            a = rng.randrange(0, len(genome) - 150)
            f.write(b"@r%d\n%s\n+\n%s\n" % (i, genome[a:a + 150], b"I" * 150))
if world > 1:
    dist.barrier()
torch.cuda.synchronize(); t0 = time.perf_counter()
text = count_fastq_sharded_sparse(path, k=21, label="x", device=local)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"rank {rank}: {dt*1e3:.1f} ms", flush=True)
if rank == 0:
    want = subprocess.run([os.path.join(ROOT, "oracle", "_build", "kpopcount_oracle"), "-k", "21", "-l", "x", "-s", path],
                          stdout=subprocess.PIPE, check=True).stdout
    print("IDENTICAL TO ORACLE" if text == want else f"MISMATCH: {len(text)} vs {len(want)} bytes", flush=True)
    os.unlink(path)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
