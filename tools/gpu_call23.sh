#!/bin/bash
# round 2, GPU call 23: bisect the synchronisation skeleton of conv_rs
mkdir -p gpurun_out
O=gpurun_out
for k in 15 31 47 79 63 127; do
  NVS_RS_KNOCK=$k timeout 300 python tools/step_breakdown.py --batch 256 > $O/c23_knock_$k.log 2>&1
  echo "== knock $k"; grep -E "^step|^ +(2|5|6|12|13) " $O/c23_knock_$k.log
done
