#!/bin/bash
# r02m: final N=1 pass of the round: smoke, whole GPU suite, bench lines, ncu launch list + full capture of the bench command.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02m_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r02m_smoke.log; tail -n 2 gpurun_out/r02m_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02m_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02m_pytest_gpu.log
grep -E "persistent|acrobot, u|comparable|passed|failed|exit|FAILED|user model" gpurun_out/r02m_pytest_gpu.log | cut -c1-250 | tail -n 30
timeout 600 python bench.py --extra > gpurun_out/r02m_bench_hopper.json 2> gpurun_out/r02m_bench_hopper.err; cut -c1-300 gpurun_out/r02m_bench_hopper.json
timeout 300 python bench.py --impl reference > gpurun_out/r02m_bench_reference.json 2>> gpurun_out/r02m_bench_hopper.err; cut -c1-200 gpurun_out/r02m_bench_reference.json
timeout 300 python bench.py --config planar_push --cpu-seconds 5 --steps 20 > gpurun_out/r02m_bench_planar_push.json 2> gpurun_out/r02m_bench_pp.err; cut -c1-300 gpurun_out/r02m_bench_planar_push.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02m_launches_bench_steps5.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contact_step_kernel -s 3 -c 1 -o gpurun_out/r02m_prof_hopper -f \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r02m_ncu_full_bench.log 2>&1; tail -n 1 gpurun_out/r02m_ncu_full_bench.log
