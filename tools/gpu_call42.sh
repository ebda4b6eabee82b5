#!/bin/bash
# round 2, GPU call 42: retrieval GEMM with clusters of 2 / 4 / 8 CTAs (database tile multicast to all of them)
mkdir -p gpurun_out
O=gpurun_out
for cs in 2 4 8; do
  NVS_RETR_CLUSTER=$cs timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q --maxfail=40 --timeout 300 > $O/c42_tests_cs$cs.log 2>&1; echo "tests cs=$cs exit $?" >> $O/c42_tests_cs$cs.log
  tail -n 2 $O/c42_tests_cs$cs.log
done
for cs in 2 4 8; do
  for stg in 3 4; do
    echo "== cluster $cs stages $stg"
    NVS_RETR_CLUSTER=$cs NVS_RETR_STAGES=$stg timeout 600 python -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c42_retr_cs${cs}_s$stg.json 2> $O/c42_retr_cs${cs}_s$stg.err; cut -c1-120 $O/c42_retr_cs${cs}_s$stg.json; grep -o '"gemm_kernel_ms": [0-9.]*' $O/c42_retr_cs${cs}_s$stg.json
  done
done
