#!/bin/bash
# (build the comparison library from a clean worktree of the commit to compare against and copy it to tools/_prev_lib.so)
# A/B on one box: tools/_prev_lib.so (library built from HEAD) against the working tree's library; C4 and the config lines
make -s -C oracle >/dev/null 2>&1
cp kpop_b200/libkpopcount_gpu.so /tmp/new.so
for round in 1 2; do
  for which in prev new; do
    if [ $which == prev ]; then cp tools/_prev_lib.so kpop_b200/libkpopcount_gpu.so; else cp /tmp/new.so kpop_b200/libkpopcount_gpu.so; fi
    echo "== $which (round $round)"
    for c in 1 2; do python bench.py --workload c4 --steps 3 --warmup 2 --genomes-per-gpu 16 --contexts $c --no-cpu 2>&1 | grep '^{' | grep -o '"ms_per_genome": [0-9.]*' | head -1 | tr '\n' ' '; done; echo
    python tools/bench_configs.py 2>/dev/null | grep -o '"config": "[A-Za-z0-9 ]*"\|"gpu_s": [0-9.]*\|"gpu_s_per_genome": [0-9.]*' | tr '\n' ' '; echo
  done
done
cp /tmp/new.so kpop_b200/libkpopcount_gpu.so
