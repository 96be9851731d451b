#!/usr/bin/env bash
# e2e leg times of the bench workload (600 frames through pinned host buffers), then per-instruction counts of the final decoder
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2ab; mkdir -p $OUT; rm -f $OUT/*
timeout 600 python scripts/e2e_time.py cfg4 32 600 > $OUT/e2e_cfg4.json 2>$OUT/err.txt; echo "e2e rc=$?"; cat $OUT/e2e_cfg4.json
timeout 300 python scripts/stage_time.py cfg4 32 2 > $OUT/plain.json 2>>$OUT/err.txt; echo "plain rc=$?"
timeout 900 ncu --clock-control none --section SourceCounters --section InstructionStats --section WarpStateStats --import-source on -k regex:bpc_decode -s 1 -c 1 -o $OUT/prof_bpc_dec python scripts/stage_time.py cfg4 32 2 > $OUT/ncu_dec.log 2>&1; echo "ncu dec rc=$?"
timeout 900 ncu --clock-control none --section SourceCounters --section InstructionStats --section WarpStateStats --import-source on -k regex:bpc_encode -s 1 -c 1 -o $OUT/prof_bpc_enc python scripts/stage_time.py cfg4 32 2 > $OUT/ncu_enc.log 2>&1; echo "ncu enc rc=$?"
tail -2 $OUT/err.txt
