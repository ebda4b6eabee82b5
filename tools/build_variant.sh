#!/bin/bash
# Build the library with one kernel file (default conv_tc.cu) taken from a git revision: tools/libnanovs_<name>.so (A/B timing on one box:
# NVS_LIB_PATH=tools/libnanovs_<name>.so python bench.py ...).  usage: tools/build_variant.sh <name> <git-rev> [file.cu]
set -e
cd "$(dirname "$0")/.."
name=$1; rev=$2; file=${3:-conv_tc.cu}
d=/tmp/nvs_variant_$name; rm -rf $d; mkdir -p $d/pkg/csrc $d/include
cp include/*.h $d/include/
cp nano_vs_slam_b200/csrc/* $d/pkg/csrc/
git show $rev:nano_vs_slam_b200/csrc/$file > $d/pkg/csrc/$file
OBJS=""
for f in $d/pkg/csrc/*.cu; do
  o=$d/$(basename ${f%.cu}).o
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -I include -c $f -o $o 2>/dev/null &
  OBJS="$OBJS $o"
done
wait
nvcc -shared -o tools/libnanovs_$name.so $OBJS
echo built tools/libnanovs_$name.so
