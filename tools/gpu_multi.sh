#!/bin/bash
# round 2, two GPUs: one-file sharding test, two devices behind one C-ABI context, C4 bench on 2 GPUs, C5-shaped bench on 2 GPUs
mkdir -p gpurun_out
make -s -C oracle > gpurun_out/oracle_build.log 2>&1
nvidia-smi -L > gpurun_out/gpus.txt
( timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -n 15 ) > gpurun_out/t_multi.log
tail -n 6 gpurun_out/t_multi.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --workload c4 --steps 3 --warmup 2 --genomes-per-gpu 16 2>&1 | tail -n 3 ) > gpurun_out/bench_c4_2gpu.log
cut -c1-1200 gpurun_out/bench_c4_2gpu.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --workload c5 --steps 3 --warmup 3 --no-e2e 2>&1 | tail -n 3 ) > gpurun_out/bench_c5_2gpu.log
cut -c1-1200 gpurun_out/bench_c5_2gpu.log
exit 0
