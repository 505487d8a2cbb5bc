#!/bin/bash
# r02u (8 GPUs): strong / weak / saturating after the 16-lane heuristic for small shards.
mkdir -p gpurun_out
run() { n=$1; tag=$2; shift 2; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus $n "$@" > gpurun_out/r02u_bench_n${n}_$tag.json 2> gpurun_out/r02u_bench_n${n}_$tag.err; echo "== n$n $tag exit $? lines $(grep -c . gpurun_out/r02u_bench_n${n}_$tag.json)"; cut -c1-260 gpurun_out/r02u_bench_n${n}_$tag.json; grep -iE "error|Traceback" gpurun_out/r02u_bench_n${n}_$tag.err | head -3; }
run 8 strong
run 4 strong
run 2 strong
