#!/bin/bash
# round 2, GPU call 15: ncu --set full of conv_rs_kernel launches (32->64, 64->64, ..., 96->64) of one V2-S step at batch 64
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_rs_kernel -s $((3*29+4)) -c 11 -o $O/r2_conv_rs -f python tools/profile_step.py --batch 64 > $O/c15_ncu.log 2>&1
ncu -i $O/r2_conv_rs.ncu-rep --page raw --csv > $O/r2_conv_rs.raw.csv 2>/dev/null
ncu -i $O/r2_conv_rs.ncu-rep --page source --csv > $O/r2_conv_rs.source.csv 2>/dev/null
tail -3 $O/c15_ncu.log
ls -la $O | grep r2_conv_rs
