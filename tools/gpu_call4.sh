#!/bin/bash
# round 2, GPU call 4: fixes of call 3 (issuer count for 3-step tiles, packed scan bound), A/B of ROW3 chunk width and
# of the retrieval kernel's stage count
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_conv_tc.py -m gpu -q --maxfail=30 > $O/c4_new_tests.log 2>&1; echo "new tests exit $?" >> $O/c4_new_tests.log
python -m pytest tests -m gpu -q --maxfail=15 > $O/c4_tests.log 2>&1; echo "tests exit $? (NVS_TC_ROW3=32)" >> $O/c4_tests.log
NVS_TC_ROW3=16 python -m pytest tests/test_gpu_model.py tests/test_torch_ops.py -m gpu -q --maxfail=15 > $O/c4_tests_row16.log 2>&1; echo "tests exit $? (NVS_TC_ROW3=16)" >> $O/c4_tests_row16.log
python tools/step_breakdown.py --batch 256 > $O/c4_breakdown_row32.log 2>&1
NVS_TC_ROW3=16 python tools/step_breakdown.py --batch 256 > $O/c4_breakdown_row16.log 2>&1
python tools/kitti_margin.py > $O/c4_kitti_margin.log 2>&1
for st in 2 3; do
  NVS_RETR_STAGES=$st python -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c4_retr_1m_s$st.json 2> $O/c4_retr_1m_s$st.err
done
NVS_TC_ROW3=16 python bench.py --steps 10 --no-retrieval --no-cpu-baseline > $O/c4_bench_row16.json 2> $O/c4_bench_row16.err
ls -la $O | grep c4_
tail -4 $O/c4_new_tests.log $O/c4_tests.log $O/c4_tests_row16.log
