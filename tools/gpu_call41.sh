#!/bin/bash
# round 2, GPU call 41: elect.sync MMA issue in the retrieval GEMM and in attention_tc
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_ops.py -m gpu -q --maxfail=40 --timeout 300 > $O/c41_tests.log 2>&1; echo "tests exit $?" >> $O/c41_tests.log
tail -n 4 $O/c41_tests.log
timeout 600 python -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c41_retr_n1.json 2> $O/c41_retr_n1.err; cat $O/c41_retr_n1.json
timeout 300 python tools/bench_attention.py > $O/c41_att.log 2>&1; cat $O/c41_att.log
