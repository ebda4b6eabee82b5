#!/bin/bash
# round 2, GPU call 52 (2 GPUs): contiguous work ranges (lists live across a cluster's whole range)
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tools/retr_waits.py 65536 1024 > $O/c52_first.log 2>&1; echo "first exit $?" >> $O/c52_first.log; tail -3 $O/c52_first.log
for cfg in "1 8" "1 4" "1 2" "0 2"; do
  set -- $cfg
  NVS_RETR_PAIR=$1 NVS_RETR_CLUSTER=$2 timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q --maxfail=40 --timeout 300 > $O/c52_tests_p$1_cs$2.log 2>&1; echo "tests pair=$1 cs=$2 exit $?" >> $O/c52_tests_p$1_cs$2.log
  tail -n 2 $O/c52_tests_p$1_cs$2.log
done
timeout 600 python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_ops.py tests/test_gpu_edges.py -m gpu -q --maxfail=40 --timeout 300 > $O/c52_tests.log 2>&1; echo "tests exit $?" >> $O/c52_tests.log; tail -n 2 $O/c52_tests.log
timeout 300 python tools/retr_waits.py 500000 10000 2>&1 | tail -10 > $O/c52_waits.log; cat $O/c52_waits.log
echo "== N=1"
timeout 600 python -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c52_retr_n1.json 2> $O/c52_retr_n1.err; grep -o '"value": [0-9.]*\|"ms_per_search": [0-9.]*\|"gemm_kernel_ms": [0-9.]*\|"achieved": [0-9.]*\|bit_exact_vs_planted": [a-z]*' $O/c52_retr_n1.json | tr '\n' ' '; echo
