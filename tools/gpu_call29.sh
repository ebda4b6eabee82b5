#!/bin/bash
# round 2, GPU call 30: conv_rs epilogue with staged TMA stores
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_rs.py -m gpu -q --maxfail=40 --timeout 120 > $O/c30_rs_tests.log 2>&1; echo "rs tests exit $?" >> $O/c30_rs_tests.log
tail -n 30 $O/c30_rs_tests.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --timeout 300 > $O/c30_tests.log 2>&1; echo "tests exit $?" >> $O/c30_tests.log
tail -n 8 $O/c30_tests.log
timeout 300 python tools/step_breakdown.py --batch 256 > $O/c30_breakdown.log 2>&1
head -32 $O/c30_breakdown.log
