#!/bin/bash
# round 2, GPU call 73: full GPU suite + smoke on the final library
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=40 --timeout 300 > $O/c73_tests.log 2>&1; echo "tests exit $?" >> $O/c73_tests.log
tail -n 3 $O/c73_tests.log
python __graft_entry__.py smoke > $O/c73_smoke.log 2>&1; tail -n 1 $O/c73_smoke.log
