#!/bin/bash
# A/B on one box: C3 through the GENERIC tile kernel (KPC_FAST=0), tools/_prev_lib.so against the working tree's library
cp kpop_b200/libkpopcount_gpu.so /tmp/new.so
for which in prev new prev new; do
  if [ $which == prev ]; then cp tools/_prev_lib.so kpop_b200/libkpopcount_gpu.so; else cp /tmp/new.so kpop_b200/libkpopcount_gpu.so; fi
  echo -n "== $which: "
  KPC_FAST=0 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --records-per-gpu 8000000 2>&1 | grep '^{' | grep -o '"ms_per_step": [0-9.]*\|"value": [0-9.e+]*' | head -2 | tr '\n' ' '; echo
done
cp /tmp/new.so kpop_b200/libkpopcount_gpu.so
