#!/bin/bash
# inner loop: CLI parity on a small input, fast-path tests, short bench, one full ncu capture
mkdir -p gpurun_out
make -s -C oracle > gpurun_out/oracle_build.log 2>&1
oracle/_build/synth_fastq 0 20000 3 > /tmp/s.fq
( timeout 120 kpop_b200/bin/KPopCount -k 12 -l x -s /tmp/s.fq | md5sum; oracle/_build/kpopcount_oracle -k 12 -l x -s /tmp/s.fq | md5sum ) > gpurun_out/cli_md5.log 2>&1
cat gpurun_out/cli_md5.log
( timeout 600 python -m pytest tests/test_gpu_fastq_fast.py -x -q 2>&1 | tail -n 30 ) > gpurun_out/t_fast.log
tail -n 4 gpurun_out/t_fast.log
( timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -n 3 ) > gpurun_out/bench_quick.log
grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"launch_ms": [0-9.]*\|"count_ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/bench_quick.log | head -n 8
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:fq_ -c 2 -o gpurun_out/fq_full -f python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --records-per-gpu 4000000 > gpurun_out/ncu_full_bench.log 2>&1 )
exit 0
