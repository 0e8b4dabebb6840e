#!/bin/bash
# full bench line (device, e2e, cpu baseline), reference arm, ncu launch list and one full capture of the fq kernels
mkdir -p gpurun_out
make -s -C oracle > gpurun_out/oracle_build.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
( timeout 900 python bench.py 2>&1 | tail -n 3 ) > gpurun_out/bench_full.log
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -n 3 ) > gpurun_out/bench_reference.log
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --records-per-gpu 8000000 > gpurun_out/ncu_launch_bench.log 2>&1 )
if [ -n "$NCU_FULL" ]; then
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:fq_ -c 2 -o gpurun_out/fq_full -f python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --records-per-gpu 4000000 > gpurun_out/ncu_full_bench.log 2>&1 )
fi
cat gpurun_out/bench_full.log gpurun_out/bench_reference.log
