#!/usr/bin/env bash
# fused forward DWT with the output base pointer pinned (kept in uniform registers): parity + timing (1 GPU)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2ai; mkdir -p $OUT; rm -f $OUT/*
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -c 200 $OUT/pytest.log
for a in "cfg4 96" "cfg4 32" "cfg2 32"; do timeout 300 python scripts/dwt_time.py $a > $OUT/dwt_$(echo $a | tr ' ' '_').json 2>$OUT/err.txt; cat $OUT/dwt_$(echo $a | tr ' ' '_').json; done
tail -2 $OUT/err.txt
