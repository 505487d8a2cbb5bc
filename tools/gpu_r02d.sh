#!/bin/bash
# r02d: GPU parity suite, bench line of every config (N=1).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d_pytest_gpu.log
tail -5 gpurun_out/r02d_pytest_gpu.log
for c in hopper acrobot cartpole_bundle planar_push rocket; do
  timeout 600 python bench.py --config $c --cpu-seconds 5 > gpurun_out/r02d_bench_$c.json 2> gpurun_out/r02d_bench_$c.err; echo "== $c exit $?"; cut -c1-400 gpurun_out/r02d_bench_$c.json; tail -3 gpurun_out/r02d_bench_$c.err
done
timeout 300 python bench.py --no-graph --no-cpu-baseline > gpurun_out/r02d_bench_hopper_eager.json 2>> gpurun_out/r02d_bench_hopper.err; cut -c1-300 gpurun_out/r02d_bench_hopper_eager.json
timeout 300 python bench.py --extra --no-cpu-baseline > gpurun_out/r02d_bench_hopper_extra.json 2>> gpurun_out/r02d_bench_hopper.err
