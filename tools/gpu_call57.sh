#!/bin/bash
# round 2, GPU call 57: NetVLAD partial kernel on FFMA2 with double-buffered chunks
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_edges.py -m gpu -q --maxfail=40 --timeout 300 > $O/c57_tests.log 2>&1; echo "tests exit $?" >> $O/c57_tests.log
tail -n 4 $O/c57_tests.log
timeout 300 python tools/step_breakdown.py --batch 256 > $O/c57_breakdown.log 2>&1
grep -E "^step|^ +(0|25) " $O/c57_breakdown.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:netvlad -c 6 python tools/step_breakdown.py --batch 256 2>&1 | grep -E "netvlad|gpu__time_duration" | head -12
