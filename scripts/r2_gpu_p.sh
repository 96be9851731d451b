#!/usr/bin/env bash
# BPC coder steps with the warp-uniform "nobody starts a codeword" skip (decoder) and deferred codeword store (encoder):
# parity (golden streams, rare coder branches, cb32, bit depths, RGB), then stage times A/B (1 GPU)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2p; mkdir -p $OUT; rm -f $OUT/*
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_coder_branches.py tests/test_gpu_cb32.py tests/test_gpu_bitdepth.py tests/test_gpu_rgb.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -c 400 $OUT/pytest.log
for w in cfg4 cfg2; do
  timeout 300 python scripts/stage_time.py $w 32 > $OUT/stage_${w}_new.json 2>$OUT/err.txt; echo "new rc=$?"; cat $OUT/stage_${w}_new.json
  PCS_LIB_PATH=$PWD/build/variants/lib_decnoskip.so timeout 300 python scripts/stage_time.py $w 32 > $OUT/stage_${w}_decnoskip.json 2>>$OUT/err.txt; echo "decnoskip rc=$?"; cat $OUT/stage_${w}_decnoskip.json
done
tail -3 $OUT/err.txt
