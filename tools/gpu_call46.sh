#!/bin/bash
# round 2, GPU call 46: knock-outs of the retrieval GEMM (1 no list work, 2 no TMEM loads, 4 no TMA)
mkdir -p gpurun_out
O=gpurun_out
for cs in 8 2; do
for kn in 0 1 2 4 6; do
  echo "== pair, cluster $cs, knock $kn"
  NVS_RETR_KNOCK=$kn NVS_RETR_CLUSTER=$cs timeout 300 python tools/retr_waits.py 500000 10000 2>&1 | tail -9 | head -6
done
done > $O/c46_knock.log 2>&1
cat $O/c46_knock.log
