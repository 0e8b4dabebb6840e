#!/bin/bash
# A/B on one box: C1 / C2 / one C4 sample through the Python binding (tools/bench_configs.py), previous library against new
make -s -C oracle >/dev/null 2>&1
cp kpop_b200/libkpopcount_gpu.so /tmp/new.so
for which in prev new prev new; do
  if [ $which == prev ]; then cp tools/_prev_lib.so kpop_b200/libkpopcount_gpu.so; else cp /tmp/new.so kpop_b200/libkpopcount_gpu.so; fi
  echo -n "== $which: "
  python tools/bench_configs.py 2>/dev/null | grep -o '"config": "[A-Za-z0-9 ]*"\|"gpu_s": [0-9.]*\|"identical_to_oracle": [a-z]*' | tr '\n' ' '; echo
done
cp /tmp/new.so kpop_b200/libkpopcount_gpu.so
