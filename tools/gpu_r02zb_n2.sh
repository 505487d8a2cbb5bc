#!/bin/bash
# Round 2, step zb (2 GPUs): planar-push shards of >= 3072 problems run the persistent sweep + a forwarding kernel under the fused
# gather; 2-rank parity of every exchange path; planar-push and hopper bench lines at N = 2.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s > gpurun_out/r02zb_pytest_n2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02zb_pytest_n2.log; tail -5 gpurun_out/r02zb_pytest_n2.log
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 "$@" > gpurun_out/r02zb_bench_n2_$tag.json 2> gpurun_out/r02zb_bench_n2_$tag.err; echo "== $tag exit $?"; cut -c1-330 gpurun_out/r02zb_bench_n2_$tag.json; grep -v "^W\|^$" gpurun_out/r02zb_bench_n2_$tag.err | tail -3; }
run planar_push --config planar_push --no-cpu-baseline
run hopper --no-cpu-baseline
