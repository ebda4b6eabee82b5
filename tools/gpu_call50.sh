#!/bin/bash
# round 2, GPU call 50 (8 GPUs): sharded retrieval at N = 8 and N = 4, pipelined over query chunks or not
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -k pipelined --timeout 300 > $O/c50_tests.log 2>&1; echo "tests exit $?" >> $O/c50_tests.log
tail -n 2 $O/c50_tests.log
for n in 8 4; do
for ch in 4096 100000; do
  echo "== N=$n chunk $ch"
  NVS_RETR_CHUNK=$ch timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c50_retr_n${n}_c$ch.json 2> $O/c50_retr_n${n}_c$ch.err; grep -o '"value": [0-9.]*\|"ms_per_search": [0-9.]*\|"gemm_kernel_ms": [0-9.]*\|"achieved": [0-9.]*\|bit_exact_vs_planted": [a-z]*' $O/c50_retr_n${n}_c$ch.json | tr '\n' ' '; echo; tail -n 1 $O/c50_retr_n${n}_c$ch.err
done
done
