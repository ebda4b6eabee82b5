#!/bin/bash
# round 2, GPU call 45: epilogue counters of the retrieval GEMM
mkdir -p gpurun_out
O=gpurun_out
for cfg in "1 8 6" "1 8 4" "1 2 6" "0 2 3"; do
  set -- $cfg
  echo "== pair $1 cluster $2 stages $3"
  NVS_RETR_PAIR=$1 NVS_RETR_CLUSTER=$2 NVS_RETR_STAGES=$3 timeout 300 python tools/retr_waits.py 500000 10000 2>&1 | tail -9
done > $O/c45_waits.log 2>&1
cat $O/c45_waits.log
