#!/bin/bash
# round 2, GPU call 25: conv_rs v2 (split-format activations, no converters, three lean MMA issuers)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_rs.py -m gpu -q --maxfail=40 --timeout 120 > $O/c25_rs_tests.log 2>&1; echo "rs tests exit $?" >> $O/c25_rs_tests.log
tail -n 25 $O/c25_rs_tests.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --timeout 300 > $O/c25_tests.log 2>&1; echo "tests exit $?" >> $O/c25_tests.log
tail -n 12 $O/c25_tests.log
timeout 300 python tools/step_breakdown.py --batch 256 > $O/c25_breakdown.log 2>&1
NVS_RS_ISSUERS=1 timeout 300 python tools/step_breakdown.py --batch 256 > $O/c25_breakdown_1issuer.log 2>&1
head -32 $O/c25_breakdown.log
head -3 $O/c25_breakdown_1issuer.log
