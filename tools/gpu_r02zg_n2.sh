#!/bin/bash
# Round 2, step zg (2 GPUs): the 2-rank parity run and the hopper bench line at N = 2 with the final layouts (row pitch of the 4-lane kernels changed after r02zb).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s > gpurun_out/r02zg_pytest_n2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02zg_pytest_n2.log; grep -c " ok" gpurun_out/r02zg_pytest_n2.log; grep -v " ok" gpurun_out/r02zg_pytest_n2.log | tail -5
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 "$@" > gpurun_out/r02zg_bench_n2_$tag.json 2> gpurun_out/r02zg_bench_n2_$tag.err; echo "== $tag exit $?"; cut -c1-330 gpurun_out/r02zg_bench_n2_$tag.json; }
run hopper --no-cpu-baseline
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --impl reference --steps 5 --warmup 1 > gpurun_out/r02zg_bench_n2_reference.json 2> gpurun_out/r02zg_bench_n2_reference.err; echo "reference arm exit $?"; wc -l gpurun_out/r02zg_bench_n2_reference.json; cut -c1-200 gpurun_out/r02zg_bench_n2_reference.json
