#!/usr/bin/env bash
# Developer helper: an alternative build of libpicsong_b200.so for A/B timing (selected at run time with PCS_LIB_PATH).
#   scripts/build_variant.sh <name> <file.cu> "<extra nvcc flags>"   -> build/variants/lib_<name>.so
set -euo pipefail
cd "$(dirname "$0")/../cuda-image-and-video-codec_b200"
name=$1; src=$2; flags=${3:-}
mkdir -p ../build/variants/obj_$name
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ftz=true -Xcompiler -fPIC,-Wall,-Wno-unused-function"
$NV $flags -c -o ../build/variants/obj_$name/$(basename $src).o csrc/$src
objs=""
for f in api.cu dwt.cu dwt_fast.cu dwt_fused.cu color.cu bpc.cu bsb.cu host_util.cpp; do
  if [ "$f" = "$src" ]; then objs="$objs ../build/variants/obj_$name/$f.o"; else objs="$objs build/$f.o"; fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../build/variants/lib_$name.so $objs
echo built build/variants/lib_$name.so
