#!/bin/bash
# round 2, GPU call 33: full test suite + full bench line (retrieval, other configs, cpu baseline) + reference arm
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --timeout 300 > $O/c33_tests.log 2>&1; echo "tests exit $?" >> $O/c33_tests.log
tail -n 4 $O/c33_tests.log
timeout 1500 python bench.py > $O/c33_bench.json 2> $O/c33_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/c33_bench_ref.json 2> $O/c33_bench_ref.err; echo "ref exit $?"
python __graft_entry__.py smoke > $O/c33_smoke.log 2>&1; tail -n 2 $O/c33_smoke.log
tail -c 600 $O/c33_bench.err
