#!/usr/bin/env bash
# Round 2, GPU call C (1 GPU): fused level-1+2 forward DWT -- parity against per-level launches, then timing A/B.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2c; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_levels or stages_against_oracle or full_size or config4" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/rc.txt
tail -c 1500 $OUT/pytest.log
for W in cfg2 cfg4; do
  for M in 1 2; do
    PCS_DWT_FUSE=$M timeout 300 python scripts/dwt_time.py $W 32 20 >> $OUT/dwt_time.jsonl 2>> $OUT/dwt_time.err
  done
done
for RP in 32 64 128 256; do PCS_FUSED_RP=$RP timeout 300 python scripts/dwt_time.py cfg2 32 20 >> $OUT/dwt_time.jsonl 2>> $OUT/dwt_time.err; done
cat $OUT/dwt_time.jsonl; tail -3 $OUT/dwt_time.err; cat $OUT/rc.txt
