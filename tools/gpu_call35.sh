#!/bin/bash
# round 2, GPU call 35 (2 GPUs): two-phase sharded retrieval: tests on one GPU, then 1M x 10k on 1 and 2 GPUs
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q --maxfail=40 --timeout 300 > $O/c35_tests.log 2>&1; echo "tests exit $?" >> $O/c35_tests.log
tail -n 6 $O/c35_tests.log
timeout 600 python -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c35_retr_n1.json 2> $O/c35_retr_n1.err; cat $O/c35_retr_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c35_retr_n2.json 2> $O/c35_retr_n2.err; cat $O/c35_retr_n2.json; tail -n 3 $O/c35_retr_n2.err
