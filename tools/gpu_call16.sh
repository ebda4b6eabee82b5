#!/bin/bash
# round 2, GPU call 16: conv_rs stage knock-outs (which stage bounds a chunk?)
mkdir -p gpurun_out
O=gpurun_out
for k in 0 1 2 4 8 9 6 13; do
  NVS_RS_KNOCK=$k timeout 300 python tools/step_breakdown.py --batch 256 > $O/c16_knock_$k.log 2>&1
  echo "== knock $k"; grep -E "^step|^ +(1|2|5|6|8|12|13) " $O/c16_knock_$k.log
done
