#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for k in 2 4 6; do
  NVS_RS_KNOCK=$k timeout 300 python tools/step_breakdown.py --batch 256 > $O/c26_knock_$k.log 2>&1
  echo "== knock $k"; grep -E "^step|^ +(1|2|5|6|8|11|12|13) " $O/c26_knock_$k.log
done
for shape in "32 64 120 160 256" "96 64 120 160 256" "16 32 240 320 256"; do
  echo "== timeline $shape"; timeout 120 python tools/rs_timeline.py $shape 2>&1 | tail -3
done
