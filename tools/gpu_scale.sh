#!/bin/bash
# weak-scaling sweep on one box: N = 1, 2, 4, 8 (run under gpurun --gpus 8)
TAG=${1:-scale}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_gpus.txt
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
  fi
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${TAG}_n$N.json") if l.startswith("{")][-1]
    print("N=$N value %.4e ms/step %.4f kernel_ms %.4f e2e %.4e" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"]))
except Exception as e:
    print("N=$N failed:", e); print(open("gpurun_out/${TAG}_n$N.err").read()[-800:])
PY
done
