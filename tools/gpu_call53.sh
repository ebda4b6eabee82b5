#!/bin/bash
# round 2, GPU call 53 (2 GPUs): ring depth 3 / 4 / 5 with the contiguous work ranges; phase timing of the sharded search
mkdir -p gpurun_out
O=gpurun_out
for stg in 5 3 4; do
  echo "== stages $stg"
  NVS_RETR_STAGES=$stg timeout 600 python -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c53_retr_s$stg.json 2> $O/c53_retr_s$stg.err; grep -o '"value": [0-9.]*\|"ms_per_search": [0-9.]*\|"gemm_kernel_ms": [0-9.]*\|"achieved": [0-9.]*\|bit_exact_vs_planted": [a-z]*' $O/c53_retr_s$stg.json | tr '\n' ' '; echo
done
NVS_RETR_STAGES=5 timeout 300 python tools/retr_waits.py 500000 10000 2>&1 | tail -10 > $O/c53_waits_s5.log; cat $O/c53_waits_s5.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/retr_tail.py 1000000 > $O/c53_tail_n2.log 2>&1; grep -A8 "^N=" $O/c53_tail_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/retr_tail.py 250000 > $O/c53_tail_n2_250k.log 2>&1; grep -A8 "^N=" $O/c53_tail_n2_250k.log
