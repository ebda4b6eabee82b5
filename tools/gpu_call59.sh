#!/bin/bash
# round 2, GPU call 59: staged PixelShuffle stores in conv_rs
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_rs.py tests/test_gpu_model.py tests/test_gpu_edges.py -m gpu -q --maxfail=40 --timeout 300 > $O/c59_tests.log 2>&1; echo "tests exit $?" >> $O/c59_tests.log
tail -n 4 $O/c59_tests.log
for ps in 1 0; do
  NVS_RS_PS=$ps timeout 300 python tools/step_breakdown.py --batch 256 > $O/c59_breakdown_ps$ps.log 2>&1
  echo "== NVS_RS_PS=$ps"; grep -E "^step|^ +(8|9|11|18|20) " $O/c59_breakdown_ps$ps.log
done
