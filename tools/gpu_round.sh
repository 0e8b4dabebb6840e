#!/bin/bash
# round-end style pass: whole GPU test-suite, smoke, full bench line, reference arm, ncu launch list
mkdir -p gpurun_out
make -s -C oracle > gpurun_out/oracle_build.log 2>&1
( timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -n 15 ) > gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3 ) > gpurun_out/smoke.log
( timeout 900 python bench.py 2>&1 | tail -n 3 ) > gpurun_out/bench_full.log
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -n 3 ) > gpurun_out/bench_reference.log
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --records-per-gpu 8000000 > gpurun_out/ncu_launch_bench.log 2>&1 )
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench_full.log | cut -c1-2500
