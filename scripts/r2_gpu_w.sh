#!/usr/bin/env bash
# encoder context words per half, decoder table in dynamic shared memory (6 CTAs per SM possible): parity, stage times, occupancy variants
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2w; mkdir -p $OUT; rm -f $OUT/*
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_coder_branches.py tests/test_gpu_cb32.py tests/test_gpu_bitdepth.py tests/test_gpu_rgb.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -c 300 $OUT/pytest.log
timeout 300 python scripts/stage_time.py cfg4 32 > $OUT/stage_cfg4_main.json 2>$OUT/err.txt; echo "main rc=$?"; cat $OUT/stage_cfg4_main.json
for v in d6 e6 e4; do
  PCS_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 300 python scripts/stage_time.py cfg4 32 > $OUT/stage_cfg4_$v.json 2>>$OUT/err.txt; echo "$v rc=$?"; cat $OUT/stage_cfg4_$v.json
done
timeout 300 python scripts/stage_time.py cfg2 32 > $OUT/stage_cfg2_main.json 2>>$OUT/err.txt; echo "main cfg2 rc=$?"; cat $OUT/stage_cfg2_main.json
tail -3 $OUT/err.txt
