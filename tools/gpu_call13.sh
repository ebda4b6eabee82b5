#!/bin/bash
# round 2, GPU call 13: conv_rs with per-warp ring slots; attention prefetch A/B
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_rs.py -m gpu -q --maxfail=40 --timeout 120 > $O/c13_rs_tests.log 2>&1; echo "rs tests exit $?" >> $O/c13_rs_tests.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --timeout 300 > $O/c13_tests.log 2>&1; echo "tests exit $?" >> $O/c13_tests.log
timeout 300 python tools/step_breakdown.py --batch 256 > $O/c13_breakdown.log 2>&1
NVS_CONV_MATH=tf32 timeout 300 python tools/step_breakdown.py --batch 256 > $O/c13_breakdown_tf32.log 2>&1
timeout 300 python tools/bench_attention.py > $O/c13_att_pf0.log 2>&1
NVS_ATT_PREFETCH=1 timeout 300 python tools/bench_attention.py > $O/c13_att_pf1.log 2>&1
timeout 600 python bench.py --steps 10 --no-retrieval --no-cpu-baseline > $O/c13_bench.json 2> $O/c13_bench.err
tail -n 8 $O/c13_rs_tests.log $O/c13_tests.log
cat $O/c13_att_pf0.log $O/c13_att_pf1.log
head -40 $O/c13_breakdown.log
