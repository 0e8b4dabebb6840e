#!/bin/bash
mkdir -p gpurun_out
make -s -C oracle > gpurun_out/oracle_build.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_fastq_fast.py tests/test_gpu_parity.py -x -q -k "not fuzz" 2>&1 | tail -n 3 ) > gpurun_out/t_fast.log
( timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -n 3 ) > gpurun_out/bench_iter.log
( KPC_FQ_LAUNCH_BYTES=2147483648 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -n 3 ) > gpurun_out/bench_iter_2g.log
tail -n 2 gpurun_out/t_fast.log
grep -o '"ms_per_step": [0-9.]*\|"launch_ms": [0-9.]*\|"feed_ms_per_step": [0-9.]*' gpurun_out/bench_iter.log gpurun_out/bench_iter_2g.log
