#!/bin/bash
# end-of-round evidence in one call: test-suite, smoke, bench lines (C3, reference arm, C4, C5 at N = 1, C1/C2), launch list,
# compute-sanitizer over the fast path, the sort path and the -L path
bash tools/gpu_round.sh
bash tools/gpu_sanitize.sh
f1=tests/golden/inputs/clusters-small.fasta.gz; gunzip -c $f1 > /tmp/c1.fa
python - <<'PY'
import sys
sys.path.insert(0, 'tools')
from bench_configs import synth_genome
open('/tmp/g0.fa', 'wb').write(synth_genome(0)[0])
PY
head -c 600000 /tmp/g0.fa > /tmp/g0s.fa
for spec in "memcheck -k 30 -l G0 -f /tmp/g0s.fa" "memcheck -k 5 -L -f /tmp/c1.fa" "racecheck -k 30 -l G0 -f /tmp/g0s.fa"; do
  set -- $spec; tool=$1; shift
  timeout 600 compute-sanitizer --tool $tool --log-file /tmp/san.log kpop_b200/bin/KPopCount "$@" > /dev/null 2>&1
  tail -n 1 /tmp/san.log | sed "s|^|[$tool $*] |" | tee -a gpurun_out/sanitizer_other_paths.log
done
( timeout 600 python tools/bench_configs.py 2>&1 ) > gpurun_out/configs.jsonl
bash tools/gpu_c4.sh > /dev/null 2>&1
( timeout 900 python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | grep "^{" | tail -n 1 ) > gpurun_out/bench_c5_1gpu.json
cut -c1-300 gpurun_out/bench_c5_1gpu.json
