#!/usr/bin/env bash
# quick A/B of the forward DWT: parity of fused vs per-level + timings (1 GPU)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2d; mkdir -p $OUT; rm -f $OUT/dwt_time.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_levels" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" > $OUT/rc.txt
tail -c 600 $OUT/pytest.log
for W in cfg2 cfg4; do
  for M in 1 2; do
    PCS_DWT_FUSE=$M timeout 300 python scripts/dwt_time.py $W 32 20 >> $OUT/dwt_time.jsonl 2>> $OUT/dwt_time.err
  done
done
cat $OUT/dwt_time.jsonl; tail -3 $OUT/dwt_time.err; cat $OUT/rc.txt
