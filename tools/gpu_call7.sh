#!/bin/bash
# round 2, GPU call 7: attention with Q.K^T on tcgen05 (attention_tc.cu): parity, then A/B timing vs the all-FFMA kernel
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --maxfail=30 -k attention > $O/c7_att_tests.log 2>&1; echo "att tests exit $?" >> $O/c7_att_tests.log
timeout 600 python -m pytest tests -m gpu -q --maxfail=15 > $O/c7_tests.log 2>&1; echo "tests exit $?" >> $O/c7_tests.log
timeout 300 python tools/bench_attention.py > $O/c7_att_tc.log 2>&1
NVS_ATT_BACKEND=ffma timeout 300 python tools/bench_attention.py > $O/c7_att_ffma.log 2>&1
timeout 600 python tools/step_breakdown.py --batch 64 --letter S_A --height 512 --width 1024 --classes 19 > $O/c7_breakdown_cfg4.log 2>&1
tail -3 $O/c7_att_tests.log $O/c7_tests.log
cat $O/c7_att_tc.log $O/c7_att_ffma.log
head -3 $O/c7_breakdown_cfg4.log
