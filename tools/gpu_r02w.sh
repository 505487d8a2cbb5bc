#!/bin/bash
# Round 2, step w: planar push — park/resume of the long-running problems, 3-D cone step lengths split over the lanes, iterate objects
# in shared memory (variant zsmem7).  Parity of the new default first, then A/B timings.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02w_ab.txt; : > $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "planar or persistent or parked or rollout" > gpurun_out/r02w_pytest_pp.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02w_pytest_pp.log
tail -6 gpurun_out/r02w_pytest_pp.log
echo "== default library (split cone step lengths), OD_PARK_ITER sweep, planar push 25600" >> $OUT
for K in 0 12 16 20 24 32; do echo "OD_PARK_ITER=$K" >> $OUT; OD_PARK_ITER=$K timeout 200 python tools/micro/kernel_time.py planar_push 25600 10 >> $OUT 2>&1; done
for V in zsmem7 nosplit; do
  echo "== variant $V" >> $OUT
  for K in 0 20; do echo "OD_PARK_ITER=$K" >> $OUT; OD_B200_LIB=$PWD/tools/micro/_ab/$V.so OD_PARK_ITER=$K timeout 200 python tools/micro/kernel_time.py planar_push 25600 10 >> $OUT 2>&1; done
done
echo "== smaller batches (per-warp kernel below 4096), default vs nosplit" >> $OUT
for B in 1024 4096; do
  timeout 200 python tools/micro/kernel_time.py planar_push $B 10 >> $OUT 2>&1
  OD_B200_LIB=$PWD/tools/micro/_ab/nosplit.so timeout 200 python tools/micro/kernel_time.py planar_push $B 10 >> $OUT 2>&1
  OD_B200_LIB=$PWD/tools/micro/_ab/zsmem7.so timeout 200 python tools/micro/kernel_time.py planar_push $B 10 >> $OUT 2>&1
done
echo "== hopper regression check" >> $OUT
timeout 200 python tools/micro/kernel_time.py hopper 4096 50 >> $OUT 2>&1
timeout 200 python tools/micro/kernel_time.py hopper 512 50 >> $OUT 2>&1
cat $OUT
