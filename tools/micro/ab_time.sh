#!/bin/bash
# A/B kernel timing of differently built libraries on one GPU box (device-resident, CUDA events):
#   tools/micro/_ab/<name>.so built with other flags / from another revision, selected through OD_B200_LIB.
# usage (under gpurun): [AB_CONFIGS="hopper 4096 8;hopper 592 8"] bash tools/micro/ab_time.sh <tag>     (config = model batch lanes, lanes 0 = auto)
TAG=${1:-ab}
PY=${AB_PYTHON:-python}
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_ab.txt
: > $OUT
IFS=';' read -ra CFGS <<< "${AB_CONFIGS:-hopper 4096 8;hopper 4096 4;hopper 592 8;hopper 262144 4;cartpole_friction 4096 0;planar_push 1024 0}"
for LIB in default $(ls tools/micro/_ab/*.so 2>/dev/null); do
  for CFG in "${CFGS[@]}"; do
    set -- $CFG
    if [ "$LIB" = default ]; then unset OD_B200_LIB; else export OD_B200_LIB=$PWD/$LIB; fi
    if [ "$3" = 0 ]; then unset OD_LANES; else export OD_LANES=$3; fi
    echo -n "$(basename $LIB) : " >> $OUT
    timeout 120 $PY tools/micro/kernel_time.py $1 $2 50 >> $OUT 2>&1 || echo "FAILED $LIB $CFG" >> $OUT
  done
done
cat $OUT
