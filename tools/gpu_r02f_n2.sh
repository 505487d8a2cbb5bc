#!/bin/bash
# r02f (2 GPUs): 2-rank parity of every multi-GPU path, bench at N=2 (strong default, weak, collective variants), DMMA experiment.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02f_topo.txt 2>&1
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s > gpurun_out/r02f_pytest_multi_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02f_pytest_multi_gpu.log
tail -60 gpurun_out/r02f_pytest_multi_gpu.log | cut -c1-200
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 "$@" > gpurun_out/r02f_bench_n2_$tag.json 2> gpurun_out/r02f_bench_n2_$tag.err; echo "== $tag exit $?"; cut -c1-330 gpurun_out/r02f_bench_n2_$tag.json; grep -v "^W\|^$" gpurun_out/r02f_bench_n2_$tag.err | tail -3; }
run strong --cpu-seconds 3
run weak --scaling weak --no-cpu-baseline --no-extra
run weak_p2p --scaling weak --collective fused-p2p --no-cpu-baseline --no-extra
run weak_launch --scaling weak --collective fused-launch-barrier --no-cpu-baseline --no-extra
run weak_nccl --scaling weak --collective nccl --no-cpu-baseline --no-extra
run planar_push --config planar_push --no-cpu-baseline --steps 20
run rocket --config rocket --no-cpu-baseline --steps 50
run bundle --config cartpole_bundle --no-cpu-baseline --steps 50
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; cut -c1-250 gpurun_out/r02f_bench_n1.json
timeout 120 tools/micro/dmma_ift > gpurun_out/r02f_dmma_ift.txt 2>&1; cat gpurun_out/r02f_dmma_ift.txt
timeout 300 ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -c 4 --csv --log-file gpurun_out/r02f_dmma_ncu.csv tools/micro/dmma_ift > /dev/null 2>&1; tail -8 gpurun_out/r02f_dmma_ncu.csv | cut -c1-250
