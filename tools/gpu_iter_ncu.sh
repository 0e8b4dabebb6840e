#!/bin/bash
# inner-loop GPU check + DRAM bytes of the first partition launch
mkdir -p gpurun_out
make -s -C oracle > gpurun_out/oracle_build.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_fastq_fast.py -x -q 2>&1 | tail -n 2 ) > gpurun_out/t_fast.log
( timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -n 3 ) > gpurun_out/bench_iter.log
( timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fq_partition -c 1 python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --records-per-gpu 4000000 2>&1 | grep "dram__\|gpu__time" ) > gpurun_out/ncu_dram.log
tail -n 1 gpurun_out/t_fast.log
grep -o '"ms_per_step": [0-9.]*\|"launch_ms": [0-9.]*' gpurun_out/bench_iter.log
cat gpurun_out/ncu_dram.log
