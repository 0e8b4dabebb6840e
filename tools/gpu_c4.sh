#!/bin/bash
# round 2: C4 bench line (N = 1) with parity inside the run, launch list of one step, fast-path tests
mkdir -p gpurun_out
make -s -C oracle > gpurun_out/oracle_build.log 2>&1
( timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --genomes-per-gpu 16 2>&1 | tail -n 4 ) > gpurun_out/bench_c4.log
cut -c1-3000 gpurun_out/bench_c4.log
( timeout 300 python -m pytest tests/test_gpu_fastq_fast.py -x -q 2>&1 | tail -n 3 ) > gpurun_out/t_fast.log
tail -n 2 gpurun_out/t_fast.log
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/c4_launches.csv python tools/c4_one.py 2 > gpurun_out/ncu_c4.log 2>&1 )
exit 0
