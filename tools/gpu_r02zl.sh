#!/bin/bash
# Round 2, step zl: rocket kernel with block barriers at the phase boundaries (OD_ROCKET_PHASED = warps per block), A/B + parity.
mkdir -p gpurun_out
OUT=gpurun_out/r02zl_rocket_phase_barriers.txt; : > $OUT
for W in 1 4 8 1 4 8; do echo "OD_ROCKET_PHASED=$W" >> $OUT; OD_ROCKET_PHASED=$W timeout 120 python tools/micro/rocket_time.py 8192 >> $OUT 2>&1; done
cat $OUT
for W in 4 8; do OD_ROCKET_PHASED=$W timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k rocket > gpurun_out/r02zl_pytest_rocket_$W.log 2>&1; echo "W=$W pytest exit $?"; tail -n 1 gpurun_out/r02zl_pytest_rocket_$W.log; done
