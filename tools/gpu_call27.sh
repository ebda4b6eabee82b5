#!/bin/bash
# round 2, GPU call 27: conv_rs epilogue: batched TMEM loads, 16 epilogue warps at Cout = 64; one issuer by default
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_rs.py -m gpu -q --maxfail=40 --timeout 120 > $O/c27_rs_tests.log 2>&1; echo "rs tests exit $?" >> $O/c27_rs_tests.log
tail -n 4 $O/c27_rs_tests.log
for k in 0 2 4; do
  NVS_RS_KNOCK=$k timeout 300 python tools/step_breakdown.py --batch 256 > $O/c27_knock_$k.log 2>&1
  echo "== knock $k"; grep -E "^step|^ +(0|1|2|5|6|8|10|11|12|13|25) " $O/c27_knock_$k.log
done
for shape in "32 64 120 160 256" "96 64 120 160 256" "16 32 240 320 256"; do
  echo "== timeline $shape"; timeout 120 python tools/rs_timeline.py $shape 2>&1 | tail -3
done
