#!/bin/bash
# round 2, GPU call 31: ncu --set full + source of conv_rs launches 5 (32->64) .. 15 (incl. 96->64) at batch 64
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_rs_kernel -s $((3*29+4)) -c 11 -o $O/r2_conv_rs2 -f python tools/profile_step.py --batch 64 > $O/c31_ncu.log 2>&1
ncu -i $O/r2_conv_rs2.ncu-rep --page raw --csv > $O/r2_conv_rs2.raw.csv 2>/dev/null
ncu -i $O/r2_conv_rs2.ncu-rep --page source --csv > $O/r2_conv_rs2.source.csv 2>/dev/null
tail -3 $O/c31_ncu.log
