#!/bin/bash
# round 2, GPU call 10: attention_tc 3 CTAs/SM (single S stage) vs 2 CTAs/SM (two S stages, prefetched TMEM loads)
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --maxfail=30 -k attention > $O/c10_att_tests.log 2>&1; echo "att tests exit $?" >> $O/c10_att_tests.log
timeout 300 python tools/bench_attention.py > $O/c10_att_ctas3.log 2>&1
NVS_ATT_CTAS=2 timeout 300 python tools/bench_attention.py > $O/c10_att_ctas2.log 2>&1
tail -n 3 $O/c10_att_tests.log
cat $O/c10_att_ctas3.log $O/c10_att_ctas2.log
