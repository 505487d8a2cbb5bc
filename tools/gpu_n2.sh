#!/bin/bash
# N-GPU comparison of the collective variants (run under gpurun --gpus N):  bash tools/gpu_n2.sh <N> <tag>
N=${1:-2}; TAG=${2:-n$N}
mkdir -p gpurun_out
P=29700
for C in fused fused-launch-barrier nccl; do
  P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline --collective $C > gpurun_out/${TAG}_$C.json 2> gpurun_out/${TAG}_$C.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${TAG}_$C.json") if l.startswith("{")][-1]
    print("N=$N %-22s value %.4e ms/step %.4f kernel_ms %.4f check %s" % ("$C", d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["gather_check_bitwise_equal_to_nccl"]))
except Exception as e:
    print("N=$N $C failed:", e); print(open("gpurun_out/${TAG}_$C.err").read()[-1500:])
PY
done
