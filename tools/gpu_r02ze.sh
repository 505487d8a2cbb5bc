#!/bin/bash
# Round 2, step ze: (a) block-phased 128-thread blocks for the 4-lane hopper kernel (OD_PHASED, A/B); (b) 4 vs 16 lanes crossover after the pitch change.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02ze_times.txt; : > $OUT
for B in 4096 8192 32768 262144; do for P in 0 2048; do echo "B=$B OD_PHASED=$P" >> $OUT; OD_PHASED=$P timeout 200 python tools/micro/kernel_time.py hopper $B 100 >> $OUT 2>&1; done; done
for B in 512 768 1024 1280 1536 2048; do for L in 4 16; do echo "B=$B OD_LANES=$L" >> $OUT; OD_LANES=$L timeout 200 python tools/micro/kernel_time.py hopper $B 200 >> $OUT 2>&1; done; done
cat $OUT
