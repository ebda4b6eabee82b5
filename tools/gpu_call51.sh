#!/bin/bash
# round 2, GPU call 51 (2 GPUs): phase timing of the sharded search
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/retr_tail.py 1000000 > $O/c51_tail_n2.log 2>&1; grep -A8 "^N=" $O/c51_tail_n2.log
# shards of an 8-GPU run (125k rows each), on 2 GPUs: 250k rows in total
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/retr_tail.py 250000 > $O/c51_tail_n2_250k.log 2>&1; grep -A8 "^N=" $O/c51_tail_n2_250k.log
nsys --version 2>/dev/null | head -1
