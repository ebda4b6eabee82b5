#!/bin/bash
# round 2, GPU call 34: adaptive pose, two-phase sharded retrieval, conv_rs torch op: tests
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_pose.py tests/test_torch_ops.py -m gpu -q --maxfail=40 --timeout 300 > $O/c34_tests.log 2>&1; echo "tests exit $?" >> $O/c34_tests.log
tail -n 40 $O/c34_tests.log
