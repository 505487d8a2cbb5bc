#!/bin/bash
# r02g: planar push block-phased vs warp-independent, bundle fit, parity, DMMA rerun.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02g_pytest_gpu.log
tail -4 gpurun_out/r02g_pytest_gpu.log
OUT=gpurun_out/r02g_pp.txt; : > $OUT
for BS in 0 1; do for B in 25600 8192; do OD_BSYNC=$BS OD_LANES=8 timeout 120 python tools/micro/kernel_time.py planar_push $B 5 >> $OUT 2>&1; done; done
OD_BSYNC=1 timeout 300 python tools/micro/pp_breakdown.py 25600 >> $OUT 2>&1
cat $OUT
timeout 300 python bench.py --config planar_push --no-cpu-baseline --steps 20 > gpurun_out/r02g_bench_planar_push.json 2> gpurun_out/r02g_bench_pp.err; cut -c1-300 gpurun_out/r02g_bench_planar_push.json
timeout 300 python bench.py --config cartpole_bundle --no-cpu-baseline > gpurun_out/r02g_bench_bundle.json 2> gpurun_out/r02g_bench_bundle.err; cut -c1-300 gpurun_out/r02g_bench_bundle.json
timeout 120 tools/micro/dmma_ift > gpurun_out/r02g_dmma_ift.txt 2>&1; cat gpurun_out/r02g_dmma_ift.txt
timeout 300 ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:update_kernel -c 12 --csv --log-file gpurun_out/r02g_dmma_ncu.csv tools/micro/dmma_ift > /dev/null 2>&1; grep "37888" gpurun_out/r02g_dmma_ncu.csv | cut -d, -f5,13- | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:contact_step_kernel -s 2 -c 1 -o gpurun_out/r02g_prof_planar_push_bsync -f \
    python tools/micro/kernel_time.py planar_push 25600 3 > gpurun_out/r02g_ncu_pp.log 2>&1; tail -1 gpurun_out/r02g_ncu_pp.log
