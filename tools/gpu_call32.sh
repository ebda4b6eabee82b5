#!/bin/bash
# round 2, GPU call 32: conv_rs store variants: auto / staged / direct 256-bit
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_rs.py -m gpu -q --maxfail=40 --timeout 120 > $O/c32_rs_tests.log 2>&1; echo "rs tests exit $?" >> $O/c32_rs_tests.log
NVS_RS_STORE=2 timeout 600 python -m pytest tests/test_gpu_conv_rs.py -m gpu -q --maxfail=40 --timeout 120 > $O/c32_rs_tests_direct.log 2>&1; echo "rs tests (direct) exit $?" >> $O/c32_rs_tests_direct.log
tail -n 3 $O/c32_rs_tests.log $O/c32_rs_tests_direct.log
for m in 0 1 2; do
  NVS_RS_STORE=$m timeout 300 python tools/step_breakdown.py --batch 256 > $O/c32_store_$m.log 2>&1
  echo "== store mode $m"; grep -E "^step|^ +(1|2|5|6|8|11|12|13) " $O/c32_store_$m.log
done
