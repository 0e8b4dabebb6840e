#!/bin/bash
# end-of-round evidence in one call: test-suite, smoke, bench lines (C3, reference arm, C4, C5 at N = 1, C1/C2), launch list,
# compute-sanitizer over the fast path, the sort path and the -L path
bash tools/gpu_round.sh
bash tools/gpu_sanitize.sh
bash tools/gpu_sanitize_other.sh
( timeout 600 python tools/bench_configs.py 2>&1 ) > gpurun_out/configs.jsonl
bash tools/gpu_c4.sh > /dev/null 2>&1
( timeout 900 python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | grep "^{" | tail -n 1 ) > gpurun_out/bench_c5_1gpu.json
cut -c1-300 gpurun_out/bench_c5_1gpu.json
