#!/bin/bash
# round 2, GPU call 20: conv_rs non-MMA pipeline: which stage bounds it (knock-outs on top of "no MMAs"), sleep / arrival variants
mkdir -p gpurun_out
O=gpurun_out
for k in 5 6 12 7; do
  NVS_RS_KNOCK=$k timeout 300 python tools/step_breakdown.py --batch 256 > $O/c20_knock_$k.log 2>&1
  echo "== knock $k"; grep -E "^step|^ +(1|2|5|6|12|13) " $O/c20_knock_$k.log
done
for v in nosleep warparrive both; do
for k in 0 15; do
  NVS_LIB_PATH=tools/libnanovs_$v.so NVS_RS_KNOCK=$k timeout 300 python tools/step_breakdown.py --batch 256 > $O/c20_${v}_$k.log 2>&1
  echo "== $v knock $k"; grep -E "^step|^ +(1|2|5|6|12|13) " $O/c20_${v}_$k.log
done
done
