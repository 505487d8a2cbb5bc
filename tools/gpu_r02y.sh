#!/bin/bash
# Round 2, step y: row pitch of the staging area with PW/2 odd (planar push 32 -> 34 doubles: no bank-conflict replays on the row moves).
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02y_times.txt; : > $OUT
for B in 1024 4096 25600 102400; do timeout 200 python tools/micro/kernel_time.py planar_push $B 10 >> $OUT 2>&1; done
timeout 200 python tools/micro/kernel_time.py cartpole_friction 4096 20 >> $OUT 2>&1
timeout 200 python tools/micro/kernel_time.py hopper 4096 50 >> $OUT 2>&1
timeout 300 python tools/micro/pp_rollout_bench.py >> $OUT 2>&1
cat $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02y_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02y_pytest_gpu.log
tail -4 gpurun_out/r02y_pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:contact_sweep_kernel -s 2 -c 2 -o gpurun_out/r02y_prof_sweep_and_resume -f \
    python tools/micro/kernel_time.py planar_push 25600 1 > gpurun_out/r02y_ncu.log 2>&1; tail -1 gpurun_out/r02y_ncu.log
