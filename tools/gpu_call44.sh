#!/bin/bash
# round 2, GPU call 44: retrieval GEMM on cta_group::2 pairs (clusters of 2 / 4 / 8), tests + wait profile + full-size bench
mkdir -p gpurun_out
O=gpurun_out
# smallest case first: a protocol bug traps after ~4 s instead of hanging
timeout 120 python tools/retr_waits.py 65536 1024 > $O/c44_first.log 2>&1; echo "first exit $?" >> $O/c44_first.log; tail -8 $O/c44_first.log
for cs in 2 4 8; do
  NVS_RETR_CLUSTER=$cs timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q --maxfail=40 --timeout 300 > $O/c44_tests_cs$cs.log 2>&1; echo "tests cs=$cs exit $?" >> $O/c44_tests_cs$cs.log
  tail -n 2 $O/c44_tests_cs$cs.log
done
for cs in 2 4 8; do
  for stg in 4 6; do
    echo "== pair, cluster $cs stages $stg"
    NVS_RETR_CLUSTER=$cs NVS_RETR_STAGES=$stg timeout 300 python tools/retr_waits.py 500000 10000 2>&1 | tail -6
  done
done > $O/c44_waits.log 2>&1
cat $O/c44_waits.log
for cs in 2 4 8; do
  echo "== pair, cluster $cs"
  NVS_RETR_CLUSTER=$cs timeout 600 python -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c44_retr_cs${cs}.json 2> $O/c44_retr_cs${cs}.err; grep -o '"value": [0-9.]*\|"gemm_kernel_ms": [0-9.]*\|"achieved": [0-9.]*\|bit_exact_vs_planted": [a-z]*' $O/c44_retr_cs${cs}.json | tr '\n' ' '; echo
done
