#!/bin/bash
# Round 2, step za: tail of the planar-push sweep on two streams (resume launch beside the IFT of the finished problems).
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02za_times.txt; : > $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "planar or persistent or parked or rollout" > gpurun_out/r02za_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02za_pytest.log; tail -3 gpurun_out/r02za_pytest.log
for B in 4096 8192 25600 102400; do for T in 0 1; do echo "B=$B OD_TAIL_OVERLAP=$T" >> $OUT; OD_TAIL_OVERLAP=$T timeout 200 python tools/micro/kernel_time.py planar_push $B 10 >> $OUT 2>&1; done; done
for B in 2048 3072; do for P in 0 256; do echo "B=$B OD_PERSIST=$P" >> $OUT; OD_PERSIST=$P timeout 200 python tools/micro/kernel_time.py planar_push $B 10 >> $OUT 2>&1; done; done
cat $OUT
timeout 600 python bench.py --config planar_push > gpurun_out/r02za_bench_n1_planar_push.json 2> gpurun_out/r02za_bench_pp.err; tail -c 1200 gpurun_out/r02za_bench_n1_planar_push.json; tail -3 gpurun_out/r02za_bench_pp.err
