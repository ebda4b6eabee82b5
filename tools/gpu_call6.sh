#!/bin/bash
# round 2, GPU call 6: ROW3 with eight epilogue warps
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests/test_gpu_conv_tc.py -m gpu -q --maxfail=30 > $O/c6_new_tests.log 2>&1; echo "new tests exit $?" >> $O/c6_new_tests.log
python -m pytest tests -m gpu -q --maxfail=15 > $O/c6_tests.log 2>&1; echo "tests exit $?" >> $O/c6_tests.log
python tools/step_breakdown.py --batch 256 > $O/c6_breakdown_auto.log 2>&1
NVS_TC_ROW3=32 python tools/step_breakdown.py --batch 256 > $O/c6_breakdown_row32.log 2>&1
python bench.py --steps 10 --no-retrieval --no-cpu-baseline > $O/c6_bench.json 2> $O/c6_bench.err
tail -3 $O/c6_new_tests.log $O/c6_tests.log
grep -E "^step|row3" $O/c6_breakdown_auto.log $O/c6_breakdown_row32.log
