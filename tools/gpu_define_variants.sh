#!/bin/bash
# builds kpc_fastq.cu with each set of -D flags in VARIANTS (separated by ';') ON the GPU box; parity + time per variant
make -s -C oracle > /dev/null 2>&1
mkdir -p gpurun_out
oracle/_build/synth_fastq 0 20000 3 > /tmp/s.fq
want=$(oracle/_build/kpopcount_oracle -k 12 -l x -s /tmp/s.fq | md5sum)
cd kpop_b200/csrc
IFS=';' read -ra VS <<< "${VARIANTS:--DFQ_APPEND_GROUP=4;-DFQ_APPEND_GROUP=1}"
for v in "${VS[@]}"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -diag-suppress 177 $v -c kpc_fastq.cu -o _build/kpc_fastq.o || exit 1
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libkpopcount_gpu.so _build/kpc_kernels.o _build/kpc_fastq.o _build/kpc_rt_cuda.o _build/kpc_engine.o _build/kpc_multi.o _build/kpc_abi.o -cudart static || exit 1
  got=$(../bin/KPopCount -k 12 -l x -s /tmp/s.fq | md5sum)
  echo "variant [$v] parity $([ "$got" == "$want" ] && echo ok || echo MISMATCH)" | tee -a ../../gpurun_out/variants.log
  ( cd ../.. && python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | grep "^{" | tail -n 1 | grep -o '"ms_per_step": [0-9.]*\|"partition_ms_per_step": [0-9.]*\|"count_ms_per_step": [0-9.]*' | tr '\n' ' '; echo ) | tee -a ../../gpurun_out/variants.log
done
