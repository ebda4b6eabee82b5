#!/bin/bash
# round 2, GPU call 36 (4 GPUs): full GPU suite on one GPU, sharded retrieval on 4 GPUs, frames bench on 4 GPUs (short)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --timeout 300 > $O/c36_tests.log 2>&1; echo "tests exit $?" >> $O/c36_tests.log
tail -n 5 $O/c36_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c36_retr_n4.json 2> $O/c36_retr_n4.err; cat $O/c36_retr_n4.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > $O/c36_bench_n4.json 2> $O/c36_bench_n4.err; tail -c 1500 $O/c36_bench_n4.json; tail -n 3 $O/c36_bench_n4.err
