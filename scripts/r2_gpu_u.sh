#!/usr/bin/env bash
# per-instruction execution counts of the BPC kernels (ncu SourceCounters) on stage_time.py cfg4 32 (ran plain in call T)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2u; mkdir -p $OUT; rm -f $OUT/*
timeout 900 ncu --clock-control none --section SourceCounters --section InstructionStats --section WarpStateStats --import-source on -k regex:bpc_decode -s 1 -c 1 -o $OUT/prof_bpc_dec python scripts/stage_time.py cfg4 32 2 > $OUT/ncu_dec.log 2>&1; echo "ncu dec rc=$?"
timeout 900 ncu --clock-control none --section SourceCounters --section InstructionStats --section WarpStateStats --import-source on -k regex:bpc_encode -s 1 -c 1 -o $OUT/prof_bpc_enc python scripts/stage_time.py cfg4 32 2 > $OUT/ncu_enc.log 2>&1; echo "ncu enc rc=$?"
ls -la $OUT
