#!/bin/bash
# r02l (2 GPUs): the whole GPU suite (incl. 2-rank paths with the one-fence barrier and the persistent planar-push sweep), bench N=2 weak / strong.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02l_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02l_pytest_gpu.log
grep -E "persistent|passed|failed|exit|FAILED" gpurun_out/r02l_pytest_gpu.log | cut -c1-220 | tail -n 12
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 "$@" > gpurun_out/r02l_bench_n2_$tag.json 2> gpurun_out/r02l_bench_n2_$tag.err; echo "== $tag exit $?"; grep "^{" gpurun_out/r02l_bench_n2_$tag.json | cut -c1-250; }
run weak --scaling weak --no-cpu-baseline --no-extra
run weak_p2p --scaling weak --collective fused-p2p --no-cpu-baseline --no-extra
run strong --no-cpu-baseline --no-extra
