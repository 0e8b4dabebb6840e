#!/bin/bash
# compute-sanitizer over the fast FASTQ pipeline (CLI, 20,000 synthetic reads = 190 tiles)
mkdir -p gpurun_out
make -s -C oracle > gpurun_out/oracle_build.log 2>&1
oracle/_build/synth_fastq 0 20000 3 > /tmp/s.fq
for tool in memcheck racecheck synccheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_$tool.full kpop_b200/bin/KPopCount -k 12 -l x -s /tmp/s.fq > /dev/null 2>&1
  ( grep -c . gpurun_out/sanitizer_$tool.full; grep -m 12 "=========" gpurun_out/sanitizer_$tool.full | cut -c1-260; tail -n 2 gpurun_out/sanitizer_$tool.full | cut -c1-260 ) > gpurun_out/sanitizer_$tool.log
  rm -f gpurun_out/sanitizer_$tool.full
  echo "== $tool"; tail -n 4 gpurun_out/sanitizer_$tool.log
done
kpop_b200/bin/KPopCount -k 12 -l x -s /tmp/s.fq | md5sum; oracle/_build/kpopcount_oracle -k 12 -l x -s /tmp/s.fq | md5sum
