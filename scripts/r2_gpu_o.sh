#!/usr/bin/env bash
# ncu --set full of the fused inverse kernel (cfg4, 32 frames); the same command ran plain in call N
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2o; mkdir -p $OUT; rm -f $OUT/*
timeout 300 python scripts/dwt_time.py cfg4 32 3 > $OUT/plain.json 2>$OUT/err.txt; echo "plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwt_inverse_l21 -s 2 -c 1 -o $OUT/prof_inv_l21 python scripts/dwt_time.py cfg4 32 3 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
tail -3 $OUT/ncu.log
