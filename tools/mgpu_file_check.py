#!/usr/bin/env python
"""Under torchrun: one FASTQ file counted by all ranks (read-chunk sharding + NCCL table reduce, kpop_b200.distributed.
count_fastq_sharded) and compared byte for byte with the oracle on rank 0.  usage: torchrun ... tools/mgpu_file_check.py [reads]"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from kpop_b200.distributed import count_fastq_sharded, shard_fastq_byte_range

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
path = "/dev/shm/kpc_mgpu_check.fq"
if rank == 0:
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    trap = b"@x\nACGTACGTACGTACGT\n+\n@+@+@+@+@+@+@+@+\n@y\nTTGCACGTACGTAAAA\n+\n+@+@+@+@+@+@+@+@\n"
    with open(path, "wb") as f:
        f.write(trap * 1000)
        subprocess.run([os.path.join(ROOT, "oracle", "_build", "synth_fastq"), "0", str(reads), "3"], stdout=f, check=True)
        f.write(trap * 1000 + b"@cut\nACGTACGTACGTAC")
if world > 1:
    dist.barrier()
a, b = shard_fastq_byte_range(path)
torch.cuda.synchronize(); t0 = time.perf_counter()
text = count_fastq_sharded(path, k=12, label="x")
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"rank {rank}: bytes [{a}, {b}) of {os.path.getsize(path)}, {dt*1e3:.1f} ms", flush=True)
if rank == 0:
    want = subprocess.run([os.path.join(ROOT, "oracle", "_build", "kpopcount_oracle"), "-k", "12", "-l", "x", "-s", path],
                          stdout=subprocess.PIPE, check=True).stdout
    print("IDENTICAL TO ORACLE" if text == want else f"MISMATCH: {len(text)} vs {len(want)} bytes", flush=True)
    os.unlink(path)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
