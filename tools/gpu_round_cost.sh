#!/bin/bash
# experiment: marginal cost of the k-mer rounds (classification + append + copy-out) = T(rounds twice) - T(rounds once)
cd kpop_b200/csrc
for v in "" "-DFQ_X_ROUND_REPS=2"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -diag-suppress 177 $v -c kpc_fastq.cu -o _build/kpc_fastq.o || exit 1
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libkpopcount_gpu.so _build/kpc_kernels.o _build/kpc_fastq.o _build/kpc_rt_cuda.o _build/kpc_engine.o _build/kpc_abi.o -cudart static || exit 1
  echo "variant [$v]"
  ( cd ../.. && python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -n 1 | grep -o '"partition_ms_per_step": [0-9.]*\|"count_ms_per_step": [0-9.]*' )
done
