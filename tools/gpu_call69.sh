#!/bin/bash
# round 2, GPU call 69 (8 GPUs): bench.py at N = 8 exactly as the driver launches it (own arm + reference arm)
mkdir -p gpurun_out
O=gpurun_out
SECONDS=0
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 > $O/c69_bench_n8.json 2> $O/c69_bench_n8.err; echo "bench exit $? after $SECONDS s"
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > $O/c69_bench_ref_n8.json 2> $O/c69_bench_ref_n8.err; echo "ref exit $? after $SECONDS s"
tail -c 300 $O/c69_bench_n8.err
cut -c1-300 $O/c69_bench_n8.json
