#!/bin/bash
# Round 2, step zn: full ncu capture of the shipped rocket kernel with the projection (128-thread blocks, barriers at the phase boundaries).
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rocket_kernel -s 26 -c 1 -o gpurun_out/r02zn_prof_rocket_phased -f \
    python tools/micro/rocket_time.py 8192 > gpurun_out/r02zn_ncu_rocket.log 2>&1; tail -n 2 gpurun_out/r02zn_ncu_rocket.log
