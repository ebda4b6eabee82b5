#!/bin/bash
# round 2, GPU call 72 (8 GPUs): sharded retrieval of the head at N = 8 / 4 / 2 / 1
mkdir -p gpurun_out
O=gpurun_out
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c72_retr_n$n.json 2> $O/c72_retr_n$n.err
  echo "N=$n $(grep -o '"value": [0-9.]*\|"ms_per_search": [0-9.]*\|"gemm_kernel_ms": [0-9.]*\|bit_exact_vs_planted": [a-z]*' $O/c72_retr_n$n.json | tr '\n' ' ')"
done
timeout 300 python -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c72_retr_n1.json 2> $O/c72_retr_n1.err
echo "N=1 $(grep -o '"value": [0-9.]*\|"ms_per_search": [0-9.]*\|"gemm_kernel_ms": [0-9.]*\|bit_exact_vs_planted": [a-z]*' $O/c72_retr_n1.json | tr '\n' ' ')"
