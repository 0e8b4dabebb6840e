#!/bin/bash
# experiment: rebuild kpc_fastq.cu with -DFQ_PHASE_CLOCKS (+ each set of flags in VARIANTS, ';' separated) on the GPU box and
# print where one warp of the partition kernel spends its cycles (see FQ_PROBE in kpc_partition.cuh for the phase numbers)
mkdir -p gpurun_out
IFS=';' read -ra VS <<< "${VARIANTS:- }"
for v in "${VS[@]}"; do
  ( cd kpop_b200/csrc
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -diag-suppress 177 -DFQ_PHASE_CLOCKS $v -c kpc_fastq.cu -o _build/kpc_fastq.o || exit 1
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libkpopcount_gpu.so _build/kpc_kernels.o _build/kpc_fastq.o _build/kpc_rt_cuda.o _build/kpc_engine.o _build/kpc_multi.o _build/kpc_abi.o -cudart static || exit 1 )
  python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/phase_clocks.log 2>&1
  echo "variant [$v]" | tee -a gpurun_out/phases.log
  grep "fq phases" gpurun_out/phase_clocks.log | tail -2 | head -1 | tee -a gpurun_out/phases.log
  grep "^{" gpurun_out/phase_clocks.log | grep -o '"partition_ms_per_step": [0-9.]*' | tee -a gpurun_out/phases.log
done
