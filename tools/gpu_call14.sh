#!/bin/bash
# round 2, GPU call 14: conv_rs ring depth / converter count A/B (what bounds a chunk?), N-letter parity with zeroed buffers
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --timeout 300 > $O/c14_tests.log 2>&1; echo "tests exit $?" >> $O/c14_tests.log
for v in nb12 nb12w6 nb6w6; do
  NVS_LIB_PATH=tools/libnanovs_$v.so timeout 300 python tools/step_breakdown.py --batch 256 > $O/c14_breakdown_$v.log 2>&1
done
timeout 300 python tools/step_breakdown.py --batch 256 > $O/c14_breakdown_base.log 2>&1
tail -n 8 $O/c14_tests.log
for v in base nb12 nb12w6 nb6w6; do echo "== $v"; grep -E "^step|^ +(1|2|5|6|8|12|13) " $O/c14_breakdown_$v.log; done
