#!/bin/bash
# rebuilds kpc_kernels.cu with each set of -D flags in VARIANTS (';' separated) on the GPU box and runs the C4 bench line
mkdir -p gpurun_out
make -s -C oracle > /dev/null 2>&1
IFS=';' read -ra VS <<< "${VARIANTS:- }"
for v in "${VS[@]}"; do
  ( cd kpop_b200/csrc
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -diag-suppress 177 $v -c kpc_kernels.cu -o _build/kpc_kernels.o || exit 1
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libkpopcount_gpu.so _build/kpc_kernels.o _build/kpc_fastq.o _build/kpc_rt_cuda.o _build/kpc_engine.o _build/kpc_multi.o _build/kpc_abi.o -cudart static || exit 1 )
  echo "variant [$v]" | tee -a gpurun_out/c4_variants.log
  python bench.py --workload c4 --steps 3 --warmup 2 --genomes-per-gpu 16 ${C4_ARGS:-} 2>&1 | grep '^{' | grep -o '"value": [0-9.e+]*\|"ms_per_genome": [0-9.]*\|"identical": [0-9]*' | head -5 | tr '\n' ' ' | tee -a gpurun_out/c4_variants.log
  echo | tee -a gpurun_out/c4_variants.log
done
