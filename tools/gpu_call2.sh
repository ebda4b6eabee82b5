#!/bin/bash
# round 2, GPU call 2: row-stationary conv kernel (ROW3) parity + timing, parity margins, ncu of the non-GEMM kernels
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests/test_gpu_conv_tc.py -m gpu -q -k "row_stationary" > $O/c2_row3_tests.log 2>&1
R3=$?
echo "row3 tests exit $R3" >> $O/c2_row3_tests.log
if [ $R3 -ne 0 ]; then export NVS_TC_ROW3=0; fi
python -m pytest tests -m gpu -q --maxfail=15 > $O/c2_tests.log 2>&1; echo "tests exit $? (NVS_TC_ROW3=${NVS_TC_ROW3:-1})" >> $O/c2_tests.log
NVS_MARGIN_REPS=6 NVS_TC_SLICE=128 python tools/golden_margins.py > $O/c2_margins_128.log 2>&1
NVS_MARGIN_REPS=6 NVS_TC_SLICE=64 python tools/golden_margins.py > $O/c2_margins_64.log 2>&1
if [ $R3 -eq 0 ]; then
  NVS_TC_ROW3=0 NVS_MARGIN_REPS=6 NVS_TC_SLICE=128 python tools/golden_margins.py > $O/c2_margins_128_norow3.log 2>&1
  python tools/step_breakdown.py --batch 256 > $O/c2_breakdown_row3.log 2>&1
  NVS_TC_ROW3=0 python tools/step_breakdown.py --batch 256 > $O/c2_breakdown_norow3.log 2>&1
  python bench.py --steps 10 --no-retrieval --no-cpu-baseline > $O/c2_bench_row3.json 2> $O/c2_bench_row3.err
  NVS_TC_ROW3=0 python bench.py --steps 10 --no-retrieval --no-cpu-baseline > $O/c2_bench_norow3.json 2> $O/c2_bench_norow3.err
fi
python tools/ncu_workload.py small > $O/c2_small_plain.log 2>&1 && \
ncu --set full --clock-control none \
    -k regex:'stem_conv|netvlad|decode_kernel|seg_argmax|select_kernel|knn2|one_to_one|pose_' \
    -o $O/r2_small -f python tools/ncu_workload.py small > $O/c2_small_ncu.log 2>&1
python tools/ncu_workload.py att > $O/c2_att_plain.log 2>&1 && \
ncu --set full --clock-control none \
    -k regex:'attention_kernel|channel_stat|dwconv3x3|conv_kernel' \
    -o $O/r2_att -f python tools/ncu_workload.py att > $O/c2_att_ncu.log 2>&1
for r in r2_small r2_att; do
  if [ -f $O/$r.ncu-rep ]; then
    ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
    sz=$(stat -c %s $O/$r.ncu-rep); if [ $sz -gt 20000000 ]; then rm -f $O/$r.ncu-rep; fi
  fi
done
ls -la $O | tail -30
tail -3 $O/c2_row3_tests.log $O/c2_tests.log
