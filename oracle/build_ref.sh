#!/usr/bin/env bash
# Test/baseline infrastructure -- NOT product code.
#
# Builds the UNMODIFIED reference (13Karl/CUDA-Image-and-Video-codec, "PICSONG") for sm_100
# from its sources where they lie under $REF (default /root/reference), into oracle/_ref/
# (git-ignored, travels to the GPU box with gpurun).  No reference source is copied or edited:
#   * -I oracle/shims supplies nvToolsExt.h / nvToolsExtCuda.h forwarders (NVTX2 headers are
#     gone in CUDA 12.9),
#   * -DCUB_DATA -include oracle/shims/cub_shim.cuh replaces the vendored CUB 1.x (does not
#     compile under CUDA 12.9) with an exact int32 inclusive scan.
# Flags follow the reference's CMakeLists.txt:9-11 (-rdc=true -use_fast_math -lineinfo,
# -fpermissive) with the arch switched from sm_75 to sm_100.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
SRC="$REF/CUDA_ImCod"
OUT="$HERE/_ref"
if [ ! -d "$SRC" ]; then
  echo "build_ref: $SRC not present (GPU box?) -- keeping prebuilt $OUT" >&2
  exit 0
fi
OBJ="$OUT/obj${REF_SUFFIX:-}"   # REF_SUFFIX=_cb32 REF_EXTRA_FLAGS=-DCBLOCK_LENGTH=32: the 64x32-codeblock build (BASELINE configs[2] sweep)
mkdir -p "$OBJ"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-std=c++17 "-gencode=arch=compute_100,code=[sm_100,compute_100]" -rdc=true
       -use_fast_math -lineinfo -w -Xcompiler -fpermissive
       -I "$HERE/shims" -DCUB_DATA -include "$HERE/shims/cub_shim.cuh" ${REF_EXTRA_FLAGS:-})
TUS=(Launcher.cu Engines/Engine.cu Engines/CodingEngine.cu Engines/DecodingEngine.cu
     BitStreamBuilder/BitStreamBuilder.cu BitStreamBuilder/BitStreamBuilder.cpp
     Image/Image.cpp IO/CommandLineParser.cpp SupportFunctions/AuxiliarFunctions.cpp)
pids=()
objs=()
for tu in "${TUS[@]}"; do
  o="$OBJ/$(echo "$tu" | tr '/.' '__').o"
  objs+=("$o")
  if [ ! -f "$o" ] || [ "$SRC/$tu" -nt "$o" ] || [ "$HERE/shims/cub_shim.cuh" -nt "$o" ]; then
    "$NVCC" "${FLAGS[@]}" -x cu -dc "$SRC/$tu" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" "-gencode=arch=compute_100,code=[sm_100,compute_100]" -rdc=true "${objs[@]}" -o "$OUT/PICSONG_ref${REF_SUFFIX:-}"
# Stage-level driver (oracle/ref_harness.cu: our code, includes the reference's headers in place,
# links the reference's own objects minus Launcher.cu's main).
"$NVCC" "${FLAGS[@]}" -I "$SRC" -x cu -dc "$HERE/ref_harness.cu" -o "$OBJ/ref_harness.o"
hobjs=()
for o in "${objs[@]}"; do case "$o" in *Launcher_cu.o) ;; *) hobjs+=("$o");; esac; done
"$NVCC" "-gencode=arch=compute_100,code=[sm_100,compute_100]" -rdc=true "$OBJ/ref_harness.o" "${hobjs[@]}" -o "$OUT/ref_harness${REF_SUFFIX:-}"
# The reference binary needs its probability tables at run time (-LUTFolder); stage the k=0
# tables beside it (data, git-ignored, travels with oracle/_ref like the binary).
for d in n1_lossless n1_lossy video_lossless video_lossy; do
  mkdir -p "$OUT/LUT/$d"
  cp -f "$SRC/LUT/$d/header.txt" "$SRC/LUT/$d/"*.txt_0 "$OUT/LUT/$d/"
  # all bit-plane files (the k-factor mode of SURVEY 8f.2 reads `*_0 .. *_15`), for oracle/run_ref_golden.py --k
  mkdir -p "$OUT/LUT_all/$d"
  cp -f "$SRC/LUT/$d/"* "$OUT/LUT_all/$d/"
done
echo "build_ref: built $OUT/PICSONG_ref${REF_SUFFIX:-}"
