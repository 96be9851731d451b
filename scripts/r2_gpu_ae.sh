#!/usr/bin/env bash
# launch-batch size A/B: per-frame stage times at 16 / 32 / 64 / 96 frames per launch (1 GPU)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2ae; mkdir -p $OUT; rm -f $OUT/*
for b in 16 32 64 96; do timeout 300 python scripts/stage_time.py cfg4 $b > $OUT/stage_cfg4_b$b.json 2>$OUT/err.txt; echo "b$b rc=$?"; cat $OUT/stage_cfg4_b$b.json; done
tail -2 $OUT/err.txt
