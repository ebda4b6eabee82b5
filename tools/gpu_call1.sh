#!/bin/bash
# round 2, GPU call 1: full GPU suite, parity margins of the tensor-core backend with 128- and 64-wide launches,
# ncu --set full of the kernels that are not the two tcgen05 contractions
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/c1_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/c1_tests.log
NVS_MARGIN_REPS=8 NVS_TC_SLICE=128 python tools/golden_margins.py > gpurun_out/c1_margins_128.log 2>&1
NVS_MARGIN_REPS=8 NVS_TC_SLICE=64 python tools/golden_margins.py > gpurun_out/c1_margins_64.log 2>&1
python tools/ncu_workload.py small > gpurun_out/c1_small_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on \
    -k regex:'stem_conv|netvlad|decode_kernel|seg_argmax|select_kernel|knn2|one_to_one|pose_' \
    -o gpurun_out/r2_small -f python tools/ncu_workload.py small > gpurun_out/c1_small_ncu.log 2>&1
python tools/ncu_workload.py att > gpurun_out/c1_att_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on \
    -k regex:'attention_kernel|channel_stat|dwconv3x3|conv_kernel' \
    -o gpurun_out/r2_att -f python tools/ncu_workload.py att > gpurun_out/c1_att_ncu.log 2>&1
tail -3 gpurun_out/c1_tests.log
