#!/bin/bash
# round 2, GPU call 67: stem kernel with batched input-window loads
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_frontend.py tests/test_gpu_edges.py -m gpu -q --maxfail=40 --timeout 300 > $O/c67_tests.log 2>&1; echo "tests exit $?" >> $O/c67_tests.log
tail -n 3 $O/c67_tests.log
timeout 300 python tools/step_breakdown.py --batch 256 > $O/c67_breakdown.log 2>&1
grep -E "^step|^ +(0|1|25) " $O/c67_breakdown.log
timeout 300 python tools/bench_stem.py 2>&1 | tail -6
