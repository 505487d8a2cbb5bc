#!/bin/bash
# Validate a differently built library on a GPU box before making its switches the default:
#   1. build  tools/micro/_ab/<tag>.so  HERE (no GPU needed):   bash tools/gpu_validate_variant.sh build extract_sm -DOD_EXTRACT_SMEM=1
#   2. run under gpurun:                                           bash tools/gpu_validate_variant.sh run extract_sm
# Step 2 runs the whole GPU parity suite against the variant (OD_B200_LIB) and then the A/B kernel timings (default vs variant).
set -u
MODE=${1:?build|run}; TAG=${2:?tag}; shift 2
SO=tools/micro/_ab/${TAG}.so
if [ "$MODE" = build ]; then
  mkdir -p tools/micro/_ab
  # all flags of the variant must be given (they REPLACE _lib.DEFAULT_DEFINES)
  python -c "import sys; from optimization_dynamics_b200 import _lib; print(_lib.build(force=True, extra_flags=sys.argv[2:], out=sys.argv[1]))" $PWD/$SO "$@" && ls -la $SO
  exit $?
fi
mkdir -p gpurun_out
OD_B200_LIB=$PWD/$SO timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
AB_CONFIGS="hopper 4096 8;hopper 4096 4;hopper 262144 4;cartpole_friction 4096 0;acrobot_impact 4096 0;planar_push 1024 0" bash tools/micro/ab_time.sh ${TAG}
OD_B200_LIB=$PWD/$SO timeout 120 python tools/micro/rocket_time.py 8192 2>&1 | tee -a gpurun_out/${TAG}_ab.txt
