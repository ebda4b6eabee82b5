#!/bin/bash
# round 2, GPU call 55: launch list of the bench command (own kernels only), timelines / knock-outs of one-chunk conv_rs layers
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:nvs:: -c 400 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs --no-retrieval > $O/c55_ncu_bench.log 2>&1; echo "ncu exit $?"
for shape in "32 32 120 160 256" "16 32 240 320 256" "64 32 120 160 256"; do
  for kn in 0 2 128; do
    echo "== timeline $shape knock $kn"; NVS_RS_KNOCK=$kn timeout 120 python tools/rs_timeline.py $shape 2>&1 | tail -3 | cut -c1-260
  done
  echo "== timeline $shape staged"; NVS_RS_STORE=1 timeout 120 python tools/rs_timeline.py $shape 2>&1 | tail -3 | cut -c1-260
done > $O/c55_timelines.log 2>&1
cat $O/c55_timelines.log
