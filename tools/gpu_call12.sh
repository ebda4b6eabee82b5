#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tools/dbg_rs.py D > $O/c12_dbg_D.log 2>&1
timeout 120 python tools/dbg_rs.py N > $O/c12_dbg_N.log 2>&1
timeout 120 python tools/dbg_rs.py S v3 > $O/c12_dbg_S3.log 2>&1
timeout 120 python tools/dbg_rs.py F > $O/c12_dbg_F.log 2>&1
timeout 120 python tools/dbg_rs.py S x 240 320 > $O/c12_dbg_S.log 2>&1
tail -n 6 $O/c12_dbg_*.log
timeout 300 python tools/step_breakdown.py --batch 256 > $O/c12_breakdown.log 2>&1
head -40 $O/c12_breakdown.log
