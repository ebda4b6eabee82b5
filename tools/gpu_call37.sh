#!/bin/bash
# round 2, GPU call 37: NaN-robust range check + automatic tf32 fallback, both conv maths vs golden
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_conv_rs.py tests/test_gpu_edges.py tests/test_gpu_frontend.py -m gpu -q --maxfail=40 --timeout 300 > $O/c37_tests.log 2>&1; echo "tests exit $?" >> $O/c37_tests.log
tail -n 30 $O/c37_tests.log
