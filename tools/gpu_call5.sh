#!/bin/bash
# round 2, GPU call 5: resident weights in the ROW3 kernel (A/B against streaming), per-layer chunk width
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_edges.py -m gpu -q --maxfail=30 > $O/c5_new_tests.log 2>&1; echo "new tests exit $?" >> $O/c5_new_tests.log
python -m pytest tests -m gpu -q --maxfail=15 > $O/c5_tests.log 2>&1; echo "tests exit $?" >> $O/c5_tests.log
python tools/step_breakdown.py --batch 256 > $O/c5_breakdown_auto.log 2>&1
NVS_TC_WRES=0 python tools/step_breakdown.py --batch 256 > $O/c5_breakdown_auto_stream.log 2>&1
NVS_TC_ROW3=32 python tools/step_breakdown.py --batch 256 > $O/c5_breakdown_row32.log 2>&1
NVS_TC_ROW3=16 python tools/step_breakdown.py --batch 256 > $O/c5_breakdown_row16.log 2>&1
python bench.py --steps 10 --no-retrieval --no-cpu-baseline > $O/c5_bench.json 2> $O/c5_bench.err
ls -la $O | grep c5_
tail -4 $O/c5_new_tests.log $O/c5_tests.log
head -3 $O/c5_breakdown_*.log
