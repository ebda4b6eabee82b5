#!/bin/bash
# round 2, GPU call 74: ncu --set full of every conv_rs launch of one forward (batch 64) on the final library
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_rs_kernel -s $((3*29)) -c 29 -o $O/r2_conv_rs_final -f python tools/profile_step.py --batch 64 > $O/c74_ncu.log 2>&1; echo "ncu exit $?"
ncu -i $O/r2_conv_rs_final.ncu-rep --page raw --csv > $O/r2_conv_rs_final.raw.csv 2>/dev/null
tail -2 $O/c74_ncu.log
