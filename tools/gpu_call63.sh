#!/bin/bash
# round 2, GPU call 63: early accumulator release, double-buffered pool exchange
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_rs.py tests/test_gpu_model.py tests/test_gpu_edges.py tests/test_gpu_frontend.py -m gpu -q --maxfail=40 --timeout 300 > $O/c63_tests.log 2>&1; echo "tests exit $?" >> $O/c63_tests.log
tail -n 4 $O/c63_tests.log
timeout 300 python tools/step_breakdown.py --batch 256 > $O/c63_breakdown.log 2>&1
grep -E "^step|^ +[0-9]+ " $O/c63_breakdown.log
for shape in "16 32 240 320 256" "32 32 120 160 256"; do
  for kn in 0 2; do
    echo "== timeline $shape knock $kn"; NVS_RS_KNOCK=$kn timeout 120 python tools/rs_timeline.py $shape 2>&1 | tail -1 | cut -c1-200
  done
done
