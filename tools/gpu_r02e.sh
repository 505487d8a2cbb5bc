#!/bin/bash
# r02e: planar push after the cooperative rank-revealing IFT: parity, time breakdown, lanes; ncu captures of every kernel family.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "planar or golden or rollout" > gpurun_out/r02e_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02e_pytest_gpu.log
tail -4 gpurun_out/r02e_pytest_gpu.log
timeout 600 python tools/micro/pp_breakdown.py 25600 > gpurun_out/r02e_pp_breakdown.txt 2>&1; cat gpurun_out/r02e_pp_breakdown.txt
for L in 4 8 16; do OD_LANES=$L timeout 120 python tools/micro/kernel_time.py planar_push 25600 5 >> gpurun_out/r02e_pp_lanes.txt 2>&1; OD_LANES=$L timeout 120 python tools/micro/kernel_time.py planar_push 1024 5 >> gpurun_out/r02e_pp_lanes.txt 2>&1; done; cat gpurun_out/r02e_pp_lanes.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:contact_step_kernel -s 2 -c 1 -o gpurun_out/r02e_prof_planar_push -f \
    python tools/micro/kernel_time.py planar_push 25600 3 > gpurun_out/r02e_ncu_pp.log 2>&1; tail -1 gpurun_out/r02e_ncu_pp.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rocket_kernel -s 2 -c 1 -o gpurun_out/r02e_prof_rocket -f \
    python tools/micro/rocket_time.py 8192 > gpurun_out/r02e_ncu_rocket.log 2>&1; tail -1 gpurun_out/r02e_ncu_rocket.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"contact_rollout_kernel|riccati_kernel" -c 4 -o gpurun_out/r02e_prof_rollout_riccati -f \
    python tools/micro/ilqr_iteration_bench.py > gpurun_out/r02e_ncu_ilqr.log 2>&1; tail -2 gpurun_out/r02e_ncu_ilqr.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"contact_step_kernel|bundle_fit" -s 4 -c 2 -o gpurun_out/r02e_prof_bundle -f \
    python bench.py --config cartpole_bundle --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r02e_ncu_bundle.log 2>&1; tail -1 gpurun_out/r02e_ncu_bundle.log
ls -la gpurun_out/*.ncu-rep
