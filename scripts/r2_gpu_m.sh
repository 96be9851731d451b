#!/usr/bin/env bash
# fused inverse DWT (levels 2+1): parity against per-level launches, then timing A/B (1 GPU)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2m; mkdir -p $OUT; rm -f $OUT/*
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_levels or golden or cfg2" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -c 1500 $OUT/pytest.log
for w in cfg2 cfg4; do
  PCS_DWT_FUSE_INV=0 timeout 300 python scripts/dwt_time.py $w > $OUT/dwt_${w}_perlevel.json 2>$OUT/err.txt; echo "rc=$?"; cat $OUT/dwt_${w}_perlevel.json
  timeout 300 python scripts/dwt_time.py $w > $OUT/dwt_${w}_fused.json 2>>$OUT/err.txt; echo "rc=$?"; cat $OUT/dwt_${w}_fused.json
done
tail -5 $OUT/err.txt
