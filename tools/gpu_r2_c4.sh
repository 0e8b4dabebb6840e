#!/bin/bash
# round 2: the sort path for large-k samples on the GPU: parity (fixtures, fuzz, longer inputs), per-genome timing, launch list
mkdir -p gpurun_out
make -s -C oracle > gpurun_out/oracle_build.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fixtures or fuzz or longer or cli" 2>&1 | tail -n 15 ) > gpurun_out/t_parity.log
tail -n 5 gpurun_out/t_parity.log
( timeout 300 python tools/c4_one.py 4 2>&1 | tail -n 6 ) > gpurun_out/c4_one.log
cat gpurun_out/c4_one.log
( KPC_SORT_PATH=0 timeout 300 python tools/c4_one.py 3 2>&1 | tail -n 3 ) > gpurun_out/c4_one_hash.log
cat gpurun_out/c4_one_hash.log
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/c4_launches.csv python tools/c4_one.py 2 > gpurun_out/ncu_c4.log 2>&1 )
exit 0
