#!/bin/bash
# round 2, GPU call 47: append + lazy-compaction epilogue of the retrieval GEMM
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tools/retr_waits.py 65536 1024 > $O/c47_first.log 2>&1; echo "first exit $?" >> $O/c47_first.log; tail -3 $O/c47_first.log
for cfg in "1 4" "1 8" "1 2" "0 2"; do
  set -- $cfg
  NVS_RETR_PAIR=$1 NVS_RETR_CLUSTER=$2 timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q --maxfail=40 --timeout 300 > $O/c47_tests_p$1_cs$2.log 2>&1; echo "tests pair=$1 cs=$2 exit $?" >> $O/c47_tests_p$1_cs$2.log
  tail -n 2 $O/c47_tests_p$1_cs$2.log
done
for cfg in "1 8 4" "1 4 4" "1 2 4" "0 2 3" "0 8 3"; do
  set -- $cfg
  echo "== pair $1 cluster $2 stages $3"
  NVS_RETR_PAIR=$1 NVS_RETR_CLUSTER=$2 NVS_RETR_STAGES=$3 timeout 300 python tools/retr_waits.py 500000 10000 2>&1 | tail -9
done > $O/c47_waits.log 2>&1
cat $O/c47_waits.log
for cfg in "1 8" "1 4" "1 2" "0 8"; do
  set -- $cfg
  echo "== pair $1, cluster $2"
  NVS_RETR_PAIR=$1 NVS_RETR_CLUSTER=$2 timeout 600 python -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c47_retr_p$1_cs$2.json 2> $O/c47_retr_p$1_cs$2.err; grep -o '"value": [0-9.]*\|"gemm_kernel_ms": [0-9.]*\|"achieved": [0-9.]*\|bit_exact_vs_planted": [a-z]*' $O/c47_retr_p$1_cs$2.json | tr '\n' ' '; echo
done
