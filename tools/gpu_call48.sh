#!/bin/bash
# round 2, GPU call 48: full GPU suite on the pair kernel, L2 hints A/B, ncu capture of the retrieval GEMM
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --timeout 300 > $O/c48_tests.log 2>&1; echo "tests exit $?" >> $O/c48_tests.log
tail -n 3 $O/c48_tests.log
for h in 0 1 2 3; do
  echo "== hint $h"
  NVS_RETR_HINT=$h timeout 600 python -m nano_vs_slam_b200.retrieval_bench 1000000 10000 > $O/c48_retr_h$h.json 2> $O/c48_retr_h$h.err; grep -o '"value": [0-9.]*\|"gemm_kernel_ms": [0-9.]*\|"achieved": [0-9.]*\|bit_exact_vs_planted": [a-z]*' $O/c48_retr_h$h.json | tr '\n' ' '; echo
done
timeout 300 python tools/ncu_retrieval.py > $O/c48_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flat_l2_topk -c 1 -o $O/r2_retr_pair -f python tools/ncu_retrieval.py > $O/c48_ncu.log 2>&1
ncu -i $O/r2_retr_pair.ncu-rep --page raw --csv > $O/r2_retr_pair.raw.csv 2>/dev/null
ncu -i $O/r2_retr_pair.ncu-rep --page details > $O/r2_retr_pair.details.txt 2>/dev/null
tail -3 $O/c48_ncu.log
