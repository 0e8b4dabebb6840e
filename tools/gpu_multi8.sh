#!/bin/bash
# --gpus 8: C3 (weak, with the end-to-end arm), C5 (ONE 100 GB stream, strong) and C4 (sample sharding) under torchrun
mkdir -p gpurun_out
make -s -C oracle > gpurun_out/oracle_build.log 2>&1
N=${NGPU:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}" 2>&1 | grep "^{" | tail -n 1; }
( timeout 900 bash -c "$(declare -f run); N=$N; run 29621 --workload c3 --steps 5 --warmup 3" ) > gpurun_out/bench_c3_${N}gpu.json
cut -c1-400 gpurun_out/bench_c3_${N}gpu.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_c3_${N}gpu.json | cut -c1-600
( timeout 900 bash -c "$(declare -f run); N=$N; run 29622 --workload c5 --steps 5 --warmup 3 --no-e2e" ) > gpurun_out/bench_c5_${N}gpu.json
cut -c1-400 gpurun_out/bench_c5_${N}gpu.json
( timeout 900 bash -c "$(declare -f run); N=$N; run 29623 --workload c4 --steps 3 --warmup 2 --genomes-per-gpu 16" ) > gpurun_out/bench_c4_${N}gpu.json
cut -c1-400 gpurun_out/bench_c4_${N}gpu.json
exit 0
