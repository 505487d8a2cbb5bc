#!/bin/bash
# Round 2, step zj: full ncu capture of the rocket kernel WITH the thrust projection (launch 26 of rocket_time.py: 23 dynamics-only launches come first).
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rocket_kernel -s 26 -c 1 -o gpurun_out/r02zj_prof_rocket_proj -f \
    python tools/micro/rocket_time.py 8192 > gpurun_out/r02zj_ncu_rocket.log 2>&1; tail -n 2 gpurun_out/r02zj_ncu_rocket.log
