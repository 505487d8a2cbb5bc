#!/bin/bash
# Round 2, step zc: row pitch of the 4-lane configurations (hopper 22 -> 28 doubles): A/B against the previous pitch, parity suite.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02zc_times.txt; : > $OUT
for rep in 1 2 3; do
  for V in default pitch0; do
    if [ $V = default ]; then unset OD_B200_LIB; else export OD_B200_LIB=$PWD/tools/micro/_ab/$V.so; fi
    echo "== $V (run $rep)" >> $OUT
    timeout 200 python tools/micro/kernel_time.py hopper 4096 200 >> $OUT 2>&1
    timeout 200 python tools/micro/kernel_time.py hopper 262144 10 >> $OUT 2>&1
  done
done
for V in default pitch0; do
  if [ $V = default ]; then unset OD_B200_LIB; else export OD_B200_LIB=$PWD/tools/micro/_ab/$V.so; fi
  echo "== $V other models" >> $OUT
  timeout 200 python tools/micro/kernel_time.py hopper 2048 200 >> $OUT 2>&1
  timeout 200 python tools/micro/kernel_time.py cartpole_friction 4096 100 >> $OUT 2>&1
  timeout 200 python tools/micro/kernel_time.py acrobot_impact 4096 100 >> $OUT 2>&1
  timeout 200 python tools/micro/rocket_time.py 8192 >> $OUT 2>&1
done
unset OD_B200_LIB
cat $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02zc_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02zc_pytest_gpu.log; tail -3 gpurun_out/r02zc_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02zc_bench_n1_hopper.json 2> gpurun_out/r02zc_bench_hopper.err; cut -c1-400 gpurun_out/r02zc_bench_n1_hopper.json
