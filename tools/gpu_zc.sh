for Z in 0 1 2; do
  OD_ZEROCOPY=$Z python bench.py --no-cpu-baseline --steps 200 > gpurun_out/zc_$Z.json 2> gpurun_out/zc_$Z.err
  python - <<PY
import json
d=json.load(open("gpurun_out/zc_$Z.json")); print("zerocopy $Z: kernel_ms %.4f e2e_ms %.4f e2e %.3e"%(d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["e2e"]["value"]))
PY
done
