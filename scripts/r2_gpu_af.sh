#!/usr/bin/env bash
# frames per launch 64 / 96: e2e leg times, then ncu --set full of the forward DWT call for the traffic figure (1 GPU)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2af; mkdir -p $OUT; rm -f $OUT/*
for b in 32 64 96; do timeout 600 python scripts/e2e_time.py cfg4 $b 600 > $OUT/e2e_cfg4_b$b.json 2>$OUT/err.txt; echo "e2e b$b rc=$?"; cat $OUT/e2e_cfg4_b$b.json; done
for b in 64 96; do
  timeout 300 python scripts/dwt_time.py cfg4 $b 3 > $OUT/dwt_plain_b$b.json 2>>$OUT/err.txt; rc=$?; echo "dwt b$b rc=$rc"; cat $OUT/dwt_plain_b$b.json
  if [ $rc -eq 0 ]; then timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwt_forward -s 8 -c 4 -o $OUT/prof_dwt_fwd_cfg4_b$b python scripts/dwt_time.py cfg4 $b 3 > $OUT/ncu_b$b.log 2>&1; echo "ncu b$b rc=$?"; fi
done
tail -2 $OUT/err.txt
