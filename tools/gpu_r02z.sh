#!/bin/bash
# Round 2, step z: resume launch with 32 lanes per parked problem; OD_PARK_ITER and OD_PERSIST (threshold of the persistent sweep) sweeps.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02z_times.txt; : > $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "planar or persistent or parked" > gpurun_out/r02z_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02z_pytest.log; tail -3 gpurun_out/r02z_pytest.log
echo "== OD_PARK_ITER sweep, 25600 problems (resume: 32 lanes)" >> $OUT
for K in 8 12 16 20 24 32; do echo "OD_PARK_ITER=$K" >> $OUT; OD_PARK_ITER=$K timeout 200 python tools/micro/kernel_time.py planar_push 25600 10 >> $OUT 2>&1; done
echo "== threshold of the persistent sweep (OD_PERSIST=n: from n problems; 0 = per-warp kernel)" >> $OUT
for B in 512 1024 2048 4096; do for P in 0 256; do echo "B=$B OD_PERSIST=$P" >> $OUT; OD_PERSIST=$P timeout 200 python tools/micro/kernel_time.py planar_push $B 10 >> $OUT 2>&1; done; done
echo "== 8 vs 16 lanes in the sweep" >> $OUT
for B in 4096 8192 25600; do for L in 8 16; do echo "B=$B OD_LANES=$L" >> $OUT; OD_LANES=$L timeout 200 python tools/micro/kernel_time.py planar_push $B 10 >> $OUT 2>&1; done; done
cat $OUT
