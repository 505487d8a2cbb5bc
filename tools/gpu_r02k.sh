#!/bin/bash
# r02k: persistent block-phased sweep for the planar push: parity, timing against the per-warp kernel, ncu.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "planar or persistent or golden or rollout" > gpurun_out/r02k_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02k_pytest_gpu.log
tail -n 5 gpurun_out/r02k_pytest_gpu.log
OUT=gpurun_out/r02k_pp.txt; : > $OUT
for P in 0 1; do for L in 8 16; do for B in 25600 8192 4096; do echo -n "persist=$P : " >> $OUT; OD_PERSIST=$P OD_LANES=$L timeout 120 python tools/micro/kernel_time.py planar_push $B 5 >> $OUT 2>&1; done; done; done
cat $OUT
timeout 300 python bench.py --config planar_push --no-cpu-baseline --steps 20 > gpurun_out/r02k_bench_planar_push.json 2> gpurun_out/r02k_bench_pp.err; cut -c1-300 gpurun_out/r02k_bench_planar_push.json; tail -n 3 gpurun_out/r02k_bench_pp.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"contact_sweep_kernel|contact_ift_kernel" -s 4 -c 2 -o gpurun_out/r02k_prof_planar_push_persistent -f \
    python tools/micro/kernel_time.py planar_push 25600 3 > gpurun_out/r02k_ncu_pp.log 2>&1; tail -n 1 gpurun_out/r02k_ncu_pp.log
