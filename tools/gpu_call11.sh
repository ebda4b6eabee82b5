#!/bin/bash
# round 2, GPU call 11: first run of the 3xFP16 row-stationary conv (conv_rs.cu) + attention_tc CTAs/SM A/B
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_rs.py -m gpu -q --maxfail=40 -x --timeout 120 > $O/c11_rs_tests.log 2>&1; echo "rs tests exit $?" >> $O/c11_rs_tests.log
timeout 600 python -m pytest tests/test_gpu_conv_rs.py -m gpu -q --maxfail=40 --timeout 120 > $O/c11_rs_tests_all.log 2>&1; echo "rs tests exit $?" >> $O/c11_rs_tests_all.log
timeout 300 python tools/bench_attention.py > $O/c11_att_ctas3.log 2>&1
NVS_ATT_CTAS=2 timeout 300 python tools/bench_attention.py > $O/c11_att_ctas2.log 2>&1
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -m gpu -q --maxfail=40 --timeout 300 > $O/c11_model_tests.log 2>&1; echo "model tests exit $?" >> $O/c11_model_tests.log
timeout 300 python tools/step_breakdown.py --batch 256 > $O/c11_breakdown.log 2>&1
tail -n 30 $O/c11_rs_tests.log
tail -n 5 $O/c11_rs_tests_all.log $O/c11_model_tests.log
cat $O/c11_att_ctas3.log $O/c11_att_ctas2.log
head -40 $O/c11_breakdown.log
