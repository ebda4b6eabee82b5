#!/bin/bash
# round 2, GPU call 61: what the per-tile floor of conv1b (16 -> 32 @240x320, pooled) and 32 -> 32 consists of
mkdir -p gpurun_out
O=gpurun_out
for shape in "16 32 240 320 256" "32 32 120 160 256"; do
  for kn in 0 2 10 6 14 66; do
    echo "== timeline $shape knock $kn"; NVS_RS_KNOCK=$kn timeout 120 python tools/rs_timeline.py $shape 2>&1 | tail -3 | cut -c1-200
  done
done > $O/c61_timelines.log 2>&1
cat $O/c61_timelines.log
