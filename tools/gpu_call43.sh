#!/bin/bash
# round 2, GPU call 43: wait profile of the retrieval GEMM roles per cluster size / ring depth
mkdir -p gpurun_out
O=gpurun_out
for cs in 2 4 8; do
  for stg in 3 4; do
    echo "== cluster $cs stages $stg"
    NVS_RETR_CLUSTER=$cs NVS_RETR_STAGES=$stg timeout 300 python tools/retr_waits.py 500000 10000 2>&1 | tail -6
  done
done > $O/c43_waits.log 2>&1
cat $O/c43_waits.log
