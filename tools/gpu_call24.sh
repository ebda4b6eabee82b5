#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for k in 127 15 0; do
  for shape in "32 32 120 160 256" "96 64 120 160 256"; do
    echo "== knock $k shape $shape"
    NVS_RS_KNOCK=$k timeout 120 python tools/rs_timeline.py $shape 2>&1 | tail -4
  done
done > $O/c24_timeline.log 2>&1
cat $O/c24_timeline.log
