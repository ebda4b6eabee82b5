#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for k in 128 132; do
  NVS_RS_KNOCK=$k timeout 300 python tools/step_breakdown.py --batch 256 > $O/c28_knock_$k.log 2>&1
  echo "== knock $k"; grep -E "^step|^ +(1|2|5|6|8|10|11|12|13) " $O/c28_knock_$k.log
done
