#!/bin/bash
# quick GPU pass: fast-path parity tests, short bench (device-resident only) at several launch sizes, ncu launch list
mkdir -p gpurun_out
make -s -C oracle > gpurun_out/oracle_build.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_fastq_fast.py -x -q 2>&1 | tail -n 30 ) > gpurun_out/t_fast.log
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "synth or truncated or longer" 2>&1 | tail -n 30 ) > gpurun_out/t_parity.log
for lb in ${LAUNCH_SIZES:-1073741824}; do
  ( KPC_FQ_LAUNCH_BYTES=$lb timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -n 3 ) > gpurun_out/bench_$lb.log
done
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --records-per-gpu 8000000 > gpurun_out/ncu_launch_bench.log 2>&1 )
if [ -n "$NCU_FULL" ]; then
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:fq_ -c 2 -o gpurun_out/fq_full -f python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --records-per-gpu 4000000 > gpurun_out/ncu_full_bench.log 2>&1 )
fi
tail -n 4 gpurun_out/t_fast.log; tail -n 4 gpurun_out/t_parity.log
for f in gpurun_out/bench_*.log; do echo $f; grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*' $f | head -n 4; done
exit 0
