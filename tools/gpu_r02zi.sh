#!/bin/bash
# Round 2, step zi: last sanity pass on the committed tree (GPU suite, default bench line) and fresh full ncu captures of the rocket and
# gradient-bundle kernels with the final row pitches (their summaries feed roofline.traffic / compute_side of those bench lines).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02zi_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02zi_pytest_gpu.log; tail -n 3 gpurun_out/r02zi_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02zi_bench_n1_hopper.json 2> gpurun_out/r02zi_bench.err; cut -c1-330 gpurun_out/r02zi_bench_n1_hopper.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rocket_kernel -s 2 -c 1 -o gpurun_out/r02zi_prof_rocket -f \
    python tools/micro/rocket_time.py 8192 > gpurun_out/r02zi_ncu_rocket.log 2>&1; tail -n 1 gpurun_out/r02zi_ncu_rocket.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"contact_step_kernel|bundle_fit" -s 4 -c 2 -o gpurun_out/r02zi_prof_bundle -f \
    python bench.py --config cartpole_bundle --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r02zi_ncu_bundle.log 2>&1; tail -n 1 gpurun_out/r02zi_ncu_bundle.log
ls -la gpurun_out/r02zi*.ncu-rep
