#!/bin/bash
# r02i: cp.async.bulk (TMA) input staging A/B, DMMA trailing-update experiment, acrobot single-call latency, default parity re-check.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02i_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02i_pytest_gpu.log; tail -3 gpurun_out/r02i_pytest_gpu.log
OD_B200_LIB=$PWD/tools/micro/_ab/tma.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "match_oracle or golden or full_size" > gpurun_out/r02i_pytest_tma.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02i_pytest_tma.log; tail -3 gpurun_out/r02i_pytest_tma.log
AB_CONFIGS="hopper 4096 4;hopper 262144 4;hopper 1048576 4" bash tools/micro/ab_time.sh r02i
timeout 120 tools/micro/dmma_ift > gpurun_out/r02i_dmma_ift.txt 2>&1; cat gpurun_out/r02i_dmma_ift.txt
timeout 300 ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:update_kernel -c 12 --csv --log-file gpurun_out/r02i_dmma_ncu.csv tools/micro/dmma_ift > /dev/null 2>&1; grep "37888" gpurun_out/r02i_dmma_ncu.csv | cut -d, -f5,13- | tail -8
timeout 600 python bench.py --config acrobot --cpu-seconds 3 > gpurun_out/r02i_bench_acrobot.json 2> gpurun_out/r02i_bench_acrobot.err; python -c "
import json; d=json.load(open('gpurun_out/r02i_bench_acrobot.json')); print(json.dumps(d['extra'], indent=1))"
OD_B200_LIB=$PWD/tools/micro/_ab/tma.so timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,sm__inst_executed_pipe_tma.sum --clock-control none -k regex:contact_step_kernel -s 2 -c 2 --csv --log-file gpurun_out/r02i_tma_ncu.csv python tools/micro/kernel_time.py hopper 262144 3 > /dev/null 2>&1; tail -6 gpurun_out/r02i_tma_ncu.csv | cut -d, -f5,13-
