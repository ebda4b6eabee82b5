#!/bin/bash
# round 2, GPU call 22: conv_rs with converter-issued row loads (no producer thread / empty barriers), MMA warp on scheduler 3
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_rs.py -m gpu -q --maxfail=40 --timeout 120 > $O/c22_rs_tests.log 2>&1; echo "rs tests exit $?" >> $O/c22_rs_tests.log
for k in 0 15 4 6; do
  NVS_RS_KNOCK=$k timeout 300 python tools/step_breakdown.py --batch 256 > $O/c22_knock_$k.log 2>&1
  echo "== knock $k"; grep -E "^step|^ +(1|2|5|6|8|10|12|13) " $O/c22_knock_$k.log
done
tail -n 3 $O/c22_rs_tests.log
