#!/bin/bash
# builds profiling variants of the partition kernel ON the GPU box and prints their phase timers
cd kpop_b200/csrc
for v in "-DFQ_PSET=1" ; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -diag-suppress 177 -DFQ_PROFILE $v -c kpc_fastq.cu -o _build/kpc_fastq.o || exit 1
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libkpopcount_gpu.so _build/kpc_kernels.o _build/kpc_fastq.o _build/kpc_rt_cuda.o _build/kpc_engine.o _build/kpc_abi.o -cudart static || exit 1
  echo "variant $v"
  ( cd ../.. && KPC_FQ_CTAS_PER_SM=1 python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --records-per-gpu 3000000 2>&1 | grep "FQPROF" | head -2; python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --records-per-gpu 3000000 2>&1 | grep "FQPROF" | head -2 )
done
