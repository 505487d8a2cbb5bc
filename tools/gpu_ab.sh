#!/bin/bash
# One GPU-box session for an A/B round: device accuracy of fastmath.cuh, GPU parity tests on the default build, kernel timings.
TAG=${1:-ab}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
[ -x tools/micro/rcp_test ] && timeout 60 tools/micro/rcp_test > gpurun_out/${TAG}_fastmath_device.txt 2>&1; cat gpurun_out/${TAG}_fastmath_device.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
bash tools/micro/ab_time.sh ${TAG}
timeout 120 python tools/micro/rocket_time.py 8192 2>&1 | tee -a gpurun_out/${TAG}_ab.txt
timeout 120 python tools/micro/rocket_time.py 1024 2>&1 | tee -a gpurun_out/${TAG}_ab.txt
