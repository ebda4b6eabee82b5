#!/bin/bash
# round 2, GPU call 8: ncu --set full of attention_tc_kernel at config-4 size (batch 2), source page for stall reasons
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -c 1 -o $O/r2_att_tc -f python tools/ncu_workload.py att > $O/c8_ncu.log 2>&1
ncu -i $O/r2_att_tc.ncu-rep --page raw --csv > $O/r2_att_tc.raw.csv 2>/dev/null
ncu -i $O/r2_att_tc.ncu-rep --page source --csv > $O/r2_att_tc.source.csv 2>/dev/null
ncu -i $O/r2_att_tc.ncu-rep --page details > $O/r2_att_tc.details.txt 2>/dev/null
tail -5 $O/c8_ncu.log
ls -la $O | grep r2_att_tc
