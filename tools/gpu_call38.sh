#!/bin/bash
# round 2, GPU call 38: 16 epilogue warps at Cout = 32 (8 channels per warp), range check + tf32 fallback, full suite
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --timeout 300 > $O/c38_tests.log 2>&1; echo "tests exit $?" >> $O/c38_tests.log
tail -n 30 $O/c38_tests.log
timeout 300 python tools/step_breakdown.py --batch 256 > $O/c38_breakdown.log 2>&1
head -32 $O/c38_breakdown.log
