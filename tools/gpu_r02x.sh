#!/bin/bash
# Round 2, step x: new planar-push default (iterate objects in shared memory, split cone step lengths, park/resume at 16 iterations):
# whole GPU suite, bench lines, per-kernel durations of the three launches, full ncu capture of the sweep and of the resume launch.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02x_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02x_pytest_gpu.log
tail -4 gpurun_out/r02x_pytest_gpu.log
OUT=gpurun_out/r02x_times.txt; : > $OUT
for B in 1024 4096 25600 102400; do timeout 200 python tools/micro/kernel_time.py planar_push $B 10 >> $OUT 2>&1; done
timeout 300 python tools/micro/pp_rollout_bench.py >> $OUT 2>&1
timeout 200 python tools/micro/kernel_time.py hopper 4096 50 >> $OUT 2>&1
cat $OUT
timeout 600 python bench.py --config planar_push > gpurun_out/r02x_bench_n1_planar_push.json 2> gpurun_out/r02x_bench_pp.err; tail -c 1500 gpurun_out/r02x_bench_n1_planar_push.json
timeout 600 python bench.py > gpurun_out/r02x_bench_n1_hopper.json 2> gpurun_out/r02x_bench_hopper.err; tail -c 600 gpurun_out/r02x_bench_n1_hopper.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"contact_sweep_kernel|contact_ift_kernel" -c 12 --csv --log-file gpurun_out/r02x_launches_planar_push.csv \
    python tools/micro/kernel_time.py planar_push 25600 1 > /dev/null 2>&1; cut -d, -f5,9,15- gpurun_out/r02x_launches_planar_push.csv | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:contact_sweep_kernel -s 2 -c 2 -o gpurun_out/r02x_prof_sweep_and_resume -f \
    python tools/micro/kernel_time.py planar_push 25600 1 > gpurun_out/r02x_ncu.log 2>&1; tail -1 gpurun_out/r02x_ncu.log
ls -la gpurun_out/*.ncu-rep
