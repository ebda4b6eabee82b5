#!/bin/bash
# round 2, GPU call 49 (2 GPUs): pipelined sharded retrieval -- tests on one GPU, torchrun N=2 with and without chunking
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q --maxfail=40 --timeout 300 > $O/c49_tests.log 2>&1; echo "tests exit $?" >> $O/c49_tests.log
tail -n 3 $O/c49_tests.log
for ch in 4096 100000 2048; do
  echo "== N=2 chunk $ch"
  NVS_RETR_CHUNK=$ch timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c49_retr_n2_c$ch.json 2> $O/c49_retr_n2_c$ch.err; grep -o '"value": [0-9.]*\|"ms_per_search": [0-9.]*\|"gemm_kernel_ms": [0-9.]*\|"achieved": [0-9.]*\|bit_exact_vs_planted": [a-z]*' $O/c49_retr_n2_c$ch.json | tr '\n' ' '; echo; tail -n 2 $O/c49_retr_n2_c$ch.err
done
echo "== N=1"
timeout 600 python -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c49_retr_n1.json 2> $O/c49_retr_n1.err; grep -o '"value": [0-9.]*\|"ms_per_search": [0-9.]*\|"gemm_kernel_ms": [0-9.]*\|"achieved": [0-9.]*\|bit_exact_vs_planted": [a-z]*' $O/c49_retr_n1.json | tr '\n' ' '; echo
# what an 8-GPU shard looks like on one GPU: 125000 rows, all queries (GEMM + local tail, no exchange)
echo "== N=1, 125000 rows"
timeout 600 python -m nano_vs_slam_b200.retrieval_bench 125000 10000 > $O/c49_retr_125k.json 2> $O/c49_retr_125k.err; grep -o '"value": [0-9.]*\|"ms_per_search": [0-9.]*\|"gemm_kernel_ms": [0-9.]*\|"achieved": [0-9.]*' $O/c49_retr_125k.json | tr '\n' ' '; echo
