#!/bin/bash
# quick GPU check: gpu tests + hopper bench for each lanes-per-problem configuration
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
for L in 1 2 4 8; do
  OD_LANES=$L timeout 300 python bench.py --no-cpu-baseline --extra --steps 100 > gpurun_out/${TAG}_bench_L$L.json 2> gpurun_out/${TAG}_bench_L$L.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_L$L.json"))
    print("lanes $L: kernel_ms %.4f value %.3e e2e %.3e sat %.3e conv %.4f" % (d["roofline"]["kernel_ms"], d["value"], d["e2e"]["value"], d["saturating_batch"]["solves_per_s_per_gpu"], d["config"]["converged_fraction"]))
except Exception as e:
    print("lanes $L failed", e); print(open("gpurun_out/${TAG}_bench_L$L.err").read()[-1500:])
PY
done
grep -E "passed|failed|Error|error" gpurun_out/${TAG}_pytest_gpu.log | tail -8
