#!/bin/bash
# round 2, GPU call 17: conv_rs synchronisation skeleton: producer sleep / per-lane arrivals
mkdir -p gpurun_out
O=gpurun_out
for v in nosleep warparrive both; do
for k in 0 13; do
  NVS_LIB_PATH=tools/libnanovs_$v.so NVS_RS_KNOCK=$k timeout 300 python tools/step_breakdown.py --batch 256 > $O/c17_${v}_$k.log 2>&1
  echo "== $v knock $k"; grep -E "^step|^ +(1|2|5|6|12|13) " $O/c17_${v}_$k.log
done
done
