#!/bin/bash
# GPU pass over the fast FASTQ pipeline: parity, sanitizer, short bench, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
make -s -C oracle > gpurun_out/oracle_build.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_fastq_fast.py -x -q 2>&1 | tail -40 ) > gpurun_out/t_fast.log
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "kat or synth or truncated or longer" 2>&1 | tail -40 ) > gpurun_out/t_parity.log
oracle/_build/synth_fastq 0 20000 3 > /tmp/s.fq
( timeout 300 compute-sanitizer --tool memcheck kpop_b200/bin/KPopCount -k 12 -l x -s /tmp/s.fq 2>&1 | tail -30 | cut -c1-300 ) > gpurun_out/sanitizer_mem.log
( timeout 300 compute-sanitizer --tool racecheck kpop_b200/bin/KPopCount -k 12 -l x -s /tmp/s.fq 2>&1 | tail -30 | cut -c1-300 ) > gpurun_out/sanitizer_race.log
( timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -5 ) > gpurun_out/bench.log
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --records-per-gpu 8000000 > gpurun_out/ncu_launch_bench.log 2>&1 )
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:fq_ -c 4 -o gpurun_out/fq_full -f python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --records-per-gpu 8000000 > gpurun_out/ncu_full_bench.log 2>&1 )
tail -5 gpurun_out/t_fast.log gpurun_out/t_parity.log gpurun_out/bench.log
tail -3 gpurun_out/sanitizer_mem.log gpurun_out/sanitizer_race.log
