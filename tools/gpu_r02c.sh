#!/bin/bash
# r02c: GPU parity suite on the shipped build, occupancy variants at large batches, bench line, full ncu capture (4 lanes).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest_gpu.log
tail -5 gpurun_out/r02c_pytest_gpu.log
AB_CONFIGS="hopper 4096 4;hopper 16384 4;hopper 65536 4;hopper 262144 4;planar_push 25600 0" bash tools/micro/ab_time.sh r02c
timeout 600 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; cat gpurun_out/r02c_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:contact_step_kernel -s 3 -c 1 -o gpurun_out/r02c_prof -f \
    python tools/micro/kernel_time.py hopper 4096 5 > gpurun_out/r02c_ncu.log 2>&1
tail -2 gpurun_out/r02c_ncu.log
