#!/bin/bash
# r02n (8 GPUs): after the one-fence barrier: 8-rank parity, bench N=8 (strong headline + weak + saturating), peer-store variant, N=4.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811 tests/multi_gpu_worker.py > gpurun_out/r02n_eight_rank_parity.txt 2>&1; echo "worker exit $?" >> gpurun_out/r02n_eight_rank_parity.txt
grep -c " ok$" gpurun_out/r02n_eight_rank_parity.txt; grep -E "FAILED|exit|Error" gpurun_out/r02n_eight_rank_parity.txt | head -5
run() { n=$1; tag=$2; shift 2; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus $n "$@" > gpurun_out/r02n_bench_n${n}_$tag.json 2> gpurun_out/r02n_bench_n${n}_$tag.err; echo "== n$n $tag exit $?"; grep "^{" gpurun_out/r02n_bench_n${n}_$tag.json | cut -c1-260; grep -iE "error|Traceback" gpurun_out/r02n_bench_n${n}_$tag.err | head -3; }
run 8 strong --cpu-seconds 3
run 8 weak_p2p --scaling weak --collective fused-p2p --no-cpu-baseline --no-extra --steps 100
run 8 weak_launch --scaling weak --collective fused-launch-barrier --no-cpu-baseline --no-extra --steps 100
run 4 strong --no-cpu-baseline
run 8 reference --impl reference --steps 5 --warmup 1
