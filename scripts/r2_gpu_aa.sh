#!/usr/bin/env bash
# decoder step with FMA-pipe forms (mul.hi shift, b * t advance): parity + stage times; fused forward DWT at 3 CTAs per SM (A/B)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2aa; mkdir -p $OUT; rm -f $OUT/*
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_coder_branches.py tests/test_gpu_cb32.py tests/test_gpu_bitdepth.py tests/test_gpu_rgb.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -c 300 $OUT/pytest.log
for w in cfg4 cfg2; do timeout 300 python scripts/stage_time.py $w 32 > $OUT/stage_${w}.json 2>$OUT/err.txt; echo "$w rc=$?"; cat $OUT/stage_${w}.json; done
for w in cfg4 cfg2; do
  timeout 300 python scripts/dwt_time.py $w > $OUT/dwt_${w}_main.json 2>>$OUT/err.txt; cat $OUT/dwt_${w}_main.json
  PCS_LIB_PATH=$PWD/build/variants/lib_fminb3.so timeout 300 python scripts/dwt_time.py $w > $OUT/dwt_${w}_fminb3.json 2>>$OUT/err.txt; cat $OUT/dwt_${w}_fminb3.json
done
tail -3 $OUT/err.txt
