#!/bin/bash
# Debug build of the library (conv_tc timeline + stage knock-outs): tools/libnanovs_dbg.so, git-ignored.
set -e
cd "$(dirname "$0")/.."
mkdir -p /tmp/nvs_dbg_obj
OBJS=""
for f in nano_vs_slam_b200/csrc/*.cu; do
  o=/tmp/nvs_dbg_obj/$(basename ${f%.cu}).o
  if [ "$(basename $f)" = conv_tc.cu ] || [ ! -f $o ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -DNVS_TC_DEBUG -I include -c $f -o $o &
  fi
  OBJS="$OBJS $o"
done
wait
nvcc -shared -o tools/libnanovs_dbg.so $OBJS
