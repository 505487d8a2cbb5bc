#!/bin/bash
# r02p: re-validation after the bench changes (single stdout line, pinned rocket e2e): whole GPU suite + every config's bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02p_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02p_pytest_gpu.log; tail -n 3 gpurun_out/r02p_pytest_gpu.log
for c in hopper acrobot cartpole_bundle planar_push rocket; do
  timeout 600 python bench.py --config $c --cpu-seconds 5 > gpurun_out/r02p_bench_$c.json 2> gpurun_out/r02p_bench_$c.err; echo "== $c exit $? lines $(grep -c . gpurun_out/r02p_bench_$c.json)"; cut -c1-200 gpurun_out/r02p_bench_$c.json
done
