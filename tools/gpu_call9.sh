#!/bin/bash
# round 2, GPU call 9: attention_tc with backoff polling + prefetched TMEM loads
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --maxfail=30 -k attention > $O/c9_att_tests.log 2>&1; echo "att tests exit $?" >> $O/c9_att_tests.log
timeout 300 python tools/bench_attention.py > $O/c9_att_tc.log 2>&1
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q --maxfail=30 > $O/c9_model_tests.log 2>&1; echo "model tests exit $?" >> $O/c9_model_tests.log
tail -n 3 $O/c9_att_tests.log $O/c9_model_tests.log
cat $O/c9_att_tc.log
