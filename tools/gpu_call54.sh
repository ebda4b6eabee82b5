#!/bin/bash
# round 2, GPU call 54: full GPU suite, full bench line, reference arm, smoke, ncu launch list of the bench command
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --timeout 300 > $O/c54_tests.log 2>&1; echo "tests exit $?" >> $O/c54_tests.log
tail -n 3 $O/c54_tests.log
timeout 1500 python bench.py > $O/c54_bench.json 2> $O/c54_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/c54_bench_ref.json 2> $O/c54_bench_ref.err; echo "ref exit $?"
python __graft_entry__.py smoke > $O/c54_smoke.log 2>&1; tail -n 2 $O/c54_smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs --no-retrieval > $O/c54_ncu_bench.log 2>&1; echo "ncu exit $?"
tail -c 400 $O/c54_bench.err
cut -c1-400 $O/c54_bench.json
