#!/bin/bash
# Round 2, step zo: phased rocket kernel, 2 / 3 / 4 warps per block (shared memory lets one 128-thread, two 96-thread or three 64-thread blocks reside per SM).
mkdir -p gpurun_out
OUT=gpurun_out/r02zo_rocket_phased_block_size.txt; : > $OUT
for W in 4 3 2 4 3 2; do echo "OD_ROCKET_PHASED=$W" >> $OUT; OD_ROCKET_PHASED=$W timeout 60 python tools/micro/rocket_time.py 8192 2>&1 | grep "proj=1" >> $OUT; done
cat $OUT
for W in 3 2; do OD_ROCKET_PHASED=$W timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rocket" > gpurun_out/r02zo_pytest_rocket_$W.log 2>&1; echo "W=$W pytest exit $?"; tail -n 1 gpurun_out/r02zo_pytest_rocket_$W.log; done
