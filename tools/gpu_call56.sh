#!/bin/bash
# round 2, GPU call 56: per-layer times with staged / direct channels-last stores forced
mkdir -p gpurun_out
O=gpurun_out
for m in 0 1 2; do
  NVS_RS_STORE=$m timeout 300 python tools/step_breakdown.py --batch 256 > $O/c56_breakdown_store$m.log 2>&1
  echo "== NVS_RS_STORE=$m"; grep -E "^step|^ +[0-9]+ " $O/c56_breakdown_store$m.log
done
