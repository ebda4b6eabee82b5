#!/bin/bash
# round 2, GPU call 21: conv_rs ring depths again (after the division fix)
mkdir -p gpurun_out
O=gpurun_out
for v in na3 na4 nb12 na3nb12; do
for k in 0 15; do
  NVS_LIB_PATH=tools/libnanovs_$v.so NVS_RS_KNOCK=$k timeout 300 python tools/step_breakdown.py --batch 256 > $O/c21_${v}_$k.log 2>&1
  echo "== $v knock $k"; grep -E "^step|^ +(1|2|5|6|12|13) " $O/c21_${v}_$k.log
done
done
