#!/bin/bash
# r02b: lanes sweep of the shipped kernel (hopper), bench line, one full ncu capture of the 4-lane kernel at 4096 problems.
mkdir -p gpurun_out
OUT=gpurun_out/r02b_lanes.txt; : > $OUT
for B in 256 512 1024 2048 4096 8192 16384; do for L in 4 8; do
  OD_LANES=$L timeout 120 python tools/micro/kernel_time.py hopper $B 50 >> $OUT 2>&1
done; done
cat $OUT
timeout 600 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; cat gpurun_out/r02b_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:contact_step_kernel -s 3 -c 1 -o gpurun_out/r02b_prof -f \
    python tools/micro/kernel_time.py hopper 4096 5 > gpurun_out/r02b_ncu.log 2>&1
tail -2 gpurun_out/r02b_ncu.log
