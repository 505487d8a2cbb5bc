#!/bin/bash
# Round 2, step zk: large-batch parity against the oracle with the final kernels (planar push through park / resume / two-stream tail).
mkdir -p gpurun_out
OUT=gpurun_out/r02zk_parity_at_large_batches_final_kernels.txt; : > $OUT
timeout 300 python tools/micro/big_batch_parity.py 25600 planar_push >> $OUT 2>&1
timeout 300 python tools/micro/big_batch_parity.py 262144 hopper >> $OUT 2>&1
timeout 200 python tools/micro/big_batch_parity.py 65536 cartpole_friction >> $OUT 2>&1
cut -c1-420 $OUT
