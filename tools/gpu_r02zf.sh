#!/bin/bash
# Round 2, step zf: N = 1 pass with the final kernels: smoke, whole GPU suite, one bench line per config + the reference arm, ncu launch
# list of the bench command, full ncu captures of the hopper kernel and of the four planar-push launches.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02zf_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r02zf_smoke.log; tail -n 2 gpurun_out/r02zf_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02zf_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02zf_pytest_gpu.log
grep -E "passed|failed|exit|FAILED" gpurun_out/r02zf_pytest_gpu.log | tail -n 3
for C in hopper acrobot cartpole_bundle planar_push rocket; do
  timeout 600 python bench.py --config $C > gpurun_out/r02zf_bench_n1_$C.json 2> gpurun_out/r02zf_bench_$C.err; cut -c1-260 gpurun_out/r02zf_bench_n1_$C.json
done
timeout 300 python bench.py --impl reference > gpurun_out/r02zf_bench_reference_arm.json 2> gpurun_out/r02zf_bench_ref.err; cut -c1-200 gpurun_out/r02zf_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02zf_launches_bench_steps5.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02zf_ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contact_step_kernel -s 3 -c 1 -o gpurun_out/r02zf_prof_hopper -f \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r02zf_ncu_full_bench.log 2>&1; tail -n 1 gpurun_out/r02zf_ncu_full_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"contact_sweep_kernel|contact_ift_kernel" -s 4 -c 4 -o gpurun_out/r02zf_prof_planar_push -f \
    python tools/micro/kernel_time.py planar_push 25600 1 > gpurun_out/r02zf_ncu_pp.log 2>&1; tail -n 1 gpurun_out/r02zf_ncu_pp.log
ls -la gpurun_out/*.ncu-rep
for T in memcheck synccheck; do
  timeout 500 compute-sanitizer --tool $T --error-exitcode 9 python tools/micro/sanitize_paths.py > gpurun_out/r02zf_compute_sanitizer_$T.txt 2>&1; echo "$T exit $?" >> gpurun_out/r02zf_compute_sanitizer_$T.txt
  tail -n 3 gpurun_out/r02zf_compute_sanitizer_$T.txt
done
