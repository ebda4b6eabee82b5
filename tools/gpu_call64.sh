#!/bin/bash
# round 2, GPU call 64: same-box A/B of the early accumulator release + double-buffered pool exchange (head) against 68f034e
mkdir -p gpurun_out
O=gpurun_out
for rep in 1 2; do
  timeout 300 python tools/step_breakdown.py --batch 256 > $O/c64_head_$rep.log 2>&1
  NVS_LIB_PATH=tools/libnanovs_fastissue.so timeout 300 python tools/step_breakdown.py --batch 256 > $O/c64_prev_$rep.log 2>&1
  echo "rep $rep head: $(grep ^step $O/c64_head_$rep.log)   prev: $(grep ^step $O/c64_prev_$rep.log)"
done
