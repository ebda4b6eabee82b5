#!/bin/bash
# round 2, GPU call 70: full GPU suite + full bench line + reference arm + smoke on the head
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --timeout 300 > $O/c70_tests.log 2>&1; echo "tests exit $?" >> $O/c70_tests.log
tail -n 3 $O/c70_tests.log
timeout 1500 python bench.py > $O/c70_bench.json 2> $O/c70_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/c70_bench_ref.json 2> $O/c70_bench_ref.err; echo "ref exit $?"
python __graft_entry__.py smoke > $O/c70_smoke.log 2>&1; tail -n 1 $O/c70_smoke.log
timeout 300 python tools/step_breakdown.py --batch 256 > $O/c70_breakdown.log 2>&1; grep ^step $O/c70_breakdown.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:nvs:: -c 400 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs --no-retrieval > $O/c70_ncu_bench.log 2>&1; echo "ncu exit $?"
cut -c1-200 $O/c70_bench.json
