#!/bin/bash
# inner-loop GPU check: fast-path parity tests + a short device-resident bench
mkdir -p gpurun_out
make -s -C oracle > gpurun_out/oracle_build.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_fastq_fast.py -x -q 2>&1 | tail -n 5 ) > gpurun_out/t_fast.log
( timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -n 3 ) > gpurun_out/bench_iter.log
tail -n 2 gpurun_out/t_fast.log
grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*' gpurun_out/bench_iter.log | head -n 4
