#!/bin/bash
# round 2, GPU call 3: new retrieval (exact screen) + ROW3 16-channel / pooled variants: parity, then timing
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_conv_tc.py -m gpu -q --maxfail=30 > $O/c3_new_tests.log 2>&1; echo "new tests exit $?" >> $O/c3_new_tests.log
python -m pytest tests -m gpu -q --maxfail=15 > $O/c3_tests.log 2>&1; echo "tests exit $? (NVS_TC_ROW3=32)" >> $O/c3_tests.log
NVS_TC_ROW3=16 python -m pytest tests/test_gpu_model.py tests/test_torch_ops.py -m gpu -q --maxfail=15 > $O/c3_tests_row16.log 2>&1; echo "tests exit $? (NVS_TC_ROW3=16)" >> $O/c3_tests_row16.log
python tools/step_breakdown.py --batch 256 > $O/c3_breakdown_row32.log 2>&1
NVS_TC_ROW3=16 python tools/step_breakdown.py --batch 256 > $O/c3_breakdown_row16.log 2>&1
python tools/kitti_margin.py > $O/c3_kitti_margin.log 2>&1
python tools/measure_tf32_peak.py > $O/c3_tf32_peak.json 2> $O/c3_tf32_peak.err
python -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c3_retr_1m.json 2> $O/c3_retr_1m.err
NVS_RETR_STAGES=3 python -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c3_retr_1m_s3.json 2> $O/c3_retr_1m_s3.err
python bench.py --steps 10 > $O/c3_bench.json 2> $O/c3_bench.err
NVS_TC_ROW3=16 python bench.py --steps 10 --no-retrieval --no-cpu-baseline --no-other-configs > $O/c3_bench_row16.json 2> $O/c3_bench_row16.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/c3_bench_ref.json 2> $O/c3_bench_ref.err
ls -la $O | tail -20
tail -4 $O/c3_new_tests.log $O/c3_tests.log $O/c3_tests_row16.log
