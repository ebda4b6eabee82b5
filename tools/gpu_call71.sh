#!/bin/bash
# round 2, GPU call 71 (2 GPUs): parallel digit pick in the selection kernels
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q --maxfail=40 --timeout 300 > $O/c71_tests.log 2>&1; echo "tests exit $?" >> $O/c71_tests.log
tail -n 2 $O/c71_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/retr_tail.py 250000 > $O/c71_tail_n2_250k.log 2>&1; grep -A8 "^N=" $O/c71_tail_n2_250k.log
timeout 600 python -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c71_retr_n1.json 2> $O/c71_retr_n1.err; grep -o '"value": [0-9.]*\|"ms_per_search": [0-9.]*\|"gemm_kernel_ms": [0-9.]*\|bit_exact_vs_planted": [a-z]*' $O/c71_retr_n1.json | tr '\n' ' '; echo
