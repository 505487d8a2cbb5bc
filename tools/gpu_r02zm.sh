#!/bin/bash
# Round 2, step zm: rocket kernel with phase barriers as the default (with the projection): whole GPU suite incl. the new 8192-problem rocket test, rocket + hopper bench lines.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02zm_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02zm_pytest_gpu.log; grep -E "rocket proj|passed|failed|exit" gpurun_out/r02zm_pytest_gpu.log | tail -n 5
timeout 300 python bench.py --config rocket > gpurun_out/r02zm_bench_n1_rocket.json 2> gpurun_out/r02zm_bench_rocket.err; cut -c1-300 gpurun_out/r02zm_bench_n1_rocket.json
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02zm_bench_n1_hopper.json 2> gpurun_out/r02zm_bench_hopper.err; cut -c1-300 gpurun_out/r02zm_bench_n1_hopper.json
