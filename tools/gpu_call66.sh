#!/bin/bash
# round 2, GPU call 66: ncu --set full with source counters of the stem kernel
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/ncu_stem.py > $O/c66_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_conv -c 1 -o $O/r2_stem -f python tools/ncu_stem.py > $O/c66_ncu.log 2>&1
ncu -i $O/r2_stem.ncu-rep --page details > $O/r2_stem.details.txt 2>/dev/null
ncu -i $O/r2_stem.ncu-rep --page source --csv > $O/r2_stem.source.csv 2>/dev/null
tail -2 $O/c66_ncu.log
