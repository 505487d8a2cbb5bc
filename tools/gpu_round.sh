#!/bin/bash
# One GPU-box session: smoke, -m gpu tests, bench (N=1), ncu launch list + one full capture of the step kernel.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --extra > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?" >> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contact_step_kernel -s 3 -c 2 -o gpurun_out/${TAG}_prof -f \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_bench.log 2>&1
tail -3 gpurun_out/${TAG}_smoke.log; tail -5 gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
# the same kernel with one warp per SM (592 problems): the stall picture of an isolated warp, for the latency analysis in DESIGN.md
timeout 300 ncu --set full --clock-control none --import-source on -k regex:contact_step_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof_iso -f \
    python tools/micro/kernel_time.py hopper 592 5 > gpurun_out/${TAG}_ncu_iso.log 2>&1
ls -la gpurun_out/ | tail -15
# fp64 FMA peak of this GPU, for the compute-side reading of the kernel's fp64-pipe utilisation
[ -x tools/micro/fp64_peak ] && timeout 60 tools/micro/fp64_peak > gpurun_out/${TAG}_fp64_peak.txt 2>&1; tail -1 gpurun_out/${TAG}_fp64_peak.txt
