#!/bin/bash
# r02h: shared-memory-resident elimination (GroupGJS): parity on the all-models build, hopper A/B, planar push A/B with / without block phasing.
mkdir -p gpurun_out
OD_B200_LIB=$PWD/tools/micro/_ab/lasm_all.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02h_pytest_lasm_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02h_pytest_lasm_all.log
tail -3 gpurun_out/r02h_pytest_lasm_all.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "planar or golden" > gpurun_out/r02h_pytest_default.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02h_pytest_default.log
tail -3 gpurun_out/r02h_pytest_default.log
AB_CONFIGS="hopper 4096 4;hopper 4096 8;hopper 1024 8;hopper 512 8;hopper 16384 4;hopper 262144 4;cartpole_friction 4096 0;acrobot_impact 4096 0" bash tools/micro/ab_time.sh r02h
OUT=gpurun_out/r02h_pp.txt; : > $OUT
for LIB in default lareg; do for BS in 0 1; do for B in 25600 1024; do
  if [ $LIB = default ]; then unset OD_B200_LIB; else export OD_B200_LIB=$PWD/tools/micro/_ab/$LIB.so; fi
  echo -n "$LIB bsync=$BS : " >> $OUT; OD_BSYNC=$BS timeout 120 python tools/micro/kernel_time.py planar_push $B 5 >> $OUT 2>&1
done; done; done
unset OD_B200_LIB
for L in 4 8 16; do echo -n "default bsync=0 lanes=$L : " >> $OUT; OD_BSYNC=0 OD_LANES=$L timeout 120 python tools/micro/kernel_time.py planar_push 25600 5 >> $OUT 2>&1; done
cat $OUT
OD_BSYNC=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:contact_step_kernel -s 2 -c 1 -o gpurun_out/r02h_prof_planar_push_lasm -f \
    python tools/micro/kernel_time.py planar_push 25600 3 > gpurun_out/r02h_ncu_pp.log 2>&1; tail -1 gpurun_out/r02h_ncu_pp.log
OD_B200_LIB=$PWD/tools/micro/_ab/lasm_all.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:contact_step_kernel -s 3 -c 1 -o gpurun_out/r02h_prof_hopper_lasm -f \
    python tools/micro/kernel_time.py hopper 4096 5 > gpurun_out/r02h_ncu_hopper.log 2>&1; tail -1 gpurun_out/r02h_ncu_hopper.log
