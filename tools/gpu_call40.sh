#!/bin/bash
# round 2, GPU call 40: incremental tile coordinates, spinning A producer; timeline of one-chunk layers
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_rs.py tests/test_gpu_model.py -m gpu -q --maxfail=40 --timeout 300 > $O/c40_tests.log 2>&1; echo "tests exit $?" >> $O/c40_tests.log
tail -n 4 $O/c40_tests.log
timeout 300 python tools/step_breakdown.py --batch 256 > $O/c40_breakdown.log 2>&1
grep -E "^step|^ +(0|1|2|5|6|8|10|11|12|13|22) " $O/c40_breakdown.log
for shape in "16 32 240 320 256" "32 32 120 160 256" "32 64 120 160 256"; do
  echo "== timeline $shape"; timeout 120 python tools/rs_timeline.py $shape 2>&1 | tail -3
  echo "== timeline knock 2 (no epilogue) $shape"; NVS_RS_KNOCK=2 timeout 120 python tools/rs_timeline.py $shape 2>&1 | tail -1
done
