#!/bin/bash
# Round 2, step zh (8 GPUs): strong / weak / saturating hopper lines and the planar-push line with the final kernels.
mkdir -p gpurun_out
run() { n=$1; tag=$2; shift 2; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus $n "$@" > gpurun_out/r02zh_bench_n${n}_$tag.json 2> gpurun_out/r02zh_bench_n${n}_$tag.err; echo "== n$n $tag exit $? lines $(grep -c . gpurun_out/r02zh_bench_n${n}_$tag.json)"; cut -c1-260 gpurun_out/r02zh_bench_n${n}_$tag.json; grep -iE "error|Traceback" gpurun_out/r02zh_bench_n${n}_$tag.err | head -3; }
run 8 hopper --no-cpu-baseline
run 8 planar_push --config planar_push --no-cpu-baseline --steps 50
run 4 hopper --no-cpu-baseline
