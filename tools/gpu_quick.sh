#!/bin/bash
# quick GPU check: gpu tests + hopper bench for each lanes-per-problem / linear-algebra configuration
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
for CFG in "4 1" "8 1" "16 1" "4 0" "1 0"; do
  set -- $CFG; L=$1; R=$2
  OD_LANES=$L OD_REG=$R timeout 300 python bench.py --no-cpu-baseline --extra --steps 100 > gpurun_out/${TAG}_bench_L${L}R${R}.json 2> gpurun_out/${TAG}_bench_L${L}R${R}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_L${L}R${R}.json"))
    print("lanes $L reg $R: kernel_ms %.4f value %.3e e2e %.3e sat %.3e conv %.4f" % (d["roofline"]["kernel_ms"], d["value"], d["e2e"]["value"], d["saturating_batch"]["solves_per_s_per_gpu"], d["config"]["converged_fraction"]))
except Exception as e:
    print("lanes $L reg $R failed", e); print(open("gpurun_out/${TAG}_bench_L${L}R${R}.err").read()[-1500:])
PY
done
grep -E "passed|failed|Error|error" gpurun_out/${TAG}_pytest_gpu.log | tail -8
