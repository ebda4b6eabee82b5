#!/bin/bash
# round 2, GPU call 18: conv_rs A-ring depth (is the chunk handshake a latency chain?)
mkdir -p gpurun_out
O=gpurun_out
for v in na3 na4; do
for k in 0 15; do
  NVS_LIB_PATH=tools/libnanovs_$v.so NVS_RS_KNOCK=$k timeout 300 python tools/step_breakdown.py --batch 256 > $O/c18_${v}_$k.log 2>&1
  echo "== $v knock $k"; grep -E "^step|^ +(1|2|5|6|12|13) " $O/c18_${v}_$k.log
done
done
NVS_RS_KNOCK=15 timeout 300 python tools/step_breakdown.py --batch 256 > $O/c18_base_15.log 2>&1
echo "== base knock 15"; grep -E "^step|^ +(1|2|5|6|12|13) " $O/c18_base_15.log
