#!/bin/bash
# group size of the two-step bucket append (FQ_AG): parity + time per variant, built on the GPU box
make -s -C oracle > /dev/null 2>&1
cd kpop_b200/csrc
for g in ${AGS:-4 8 16}; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -diag-suppress 177 -DFQ_AG_CFG=$g -c kpc_fastq.cu -o _build/kpc_fastq.o || exit 1
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libkpopcount_gpu.so _build/kpc_kernels.o _build/kpc_fastq.o _build/kpc_rt_cuda.o _build/kpc_engine.o _build/kpc_abi.o -cudart static || exit 1
  echo "variant AG=$g"
  ( cd ../.. && python -m pytest tests/test_gpu_fastq_fast.py -x -q 2>&1 | tail -n 1; python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -n 1 | grep -o '"ms_per_step": [0-9.]*\|"partition_ms_per_step": [0-9.]*' )
done
