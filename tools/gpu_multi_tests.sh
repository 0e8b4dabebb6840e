#!/bin/bash
# round 2, two GPUs: multi-GPU tests (dense reduce, sparse merge, two devices behind one context), full-size C3 parity, fuzz
mkdir -p gpurun_out
make -s -C oracle > gpurun_out/oracle_build.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -n 15 ) > gpurun_out/t_multi.log
tail -n 4 gpurun_out/t_multi.log
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "full_size or fuzz" 2>&1 | tail -n 15 ) > gpurun_out/t_parity2.log
tail -n 4 gpurun_out/t_parity2.log
exit 0
