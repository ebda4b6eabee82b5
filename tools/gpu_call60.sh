#!/bin/bash
# round 2, GPU call 60: two epilogue warp sets for Cout = 32 launches of conv_rs
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_rs.py tests/test_gpu_model.py tests/test_gpu_edges.py tests/test_gpu_frontend.py -m gpu -q --maxfail=40 --timeout 300 > $O/c60_tests.log 2>&1; echo "tests exit $?" >> $O/c60_tests.log
tail -n 6 $O/c60_tests.log
timeout 300 python tools/step_breakdown.py --batch 256 > $O/c60_breakdown_sets.log 2>&1
echo "== two sets"; grep -E "^step|^ +(1|2|3|4|5|10|13|22) " $O/c60_breakdown_sets.log
NVS_LIB_PATH=tools/libnanovs_onesets.so timeout 300 python tools/step_breakdown.py --batch 256 > $O/c60_breakdown_one.log 2>&1
echo "== one set (1a6faa3)"; grep -E "^step|^ +(1|2|3|4|5|10|13|22) " $O/c60_breakdown_one.log
