#!/bin/bash
# one full ncu capture of the fast FASTQ kernels (first 1 GiB launch) + per-line export
mkdir -p gpurun_out
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:fq_ -c 2 -o gpurun_out/fq_full -f python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --records-per-gpu 4000000 > gpurun_out/ncu_full_bench.log 2>&1 )
tail -2 gpurun_out/ncu_full_bench.log
