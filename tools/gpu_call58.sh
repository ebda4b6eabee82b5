#!/bin/bash
# round 2, GPU call 58 (2 GPUs): bench.py at N = 2 exactly as the driver launches it
mkdir -p gpurun_out
O=gpurun_out
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 > $O/c58_bench_n2.json 2> $O/c58_bench_n2.err; echo "bench exit $? after $SECONDS s"
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/c58_bench_ref_n2.json 2> $O/c58_bench_ref_n2.err; echo "ref exit $? after $SECONDS s"
tail -c 300 $O/c58_bench_n2.err
cut -c1-300 $O/c58_bench_n2.json
