#!/usr/bin/env bash
# forward DWT: parity, A/B timing, then an ncu capture of the fused kernel (1 GPU)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2e; mkdir -p $OUT; rm -f $OUT/dwt_time.jsonl $OUT/*.ncu-rep
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_levels" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" > $OUT/rc.txt
tail -c 300 $OUT/pytest.log
for W in cfg2 cfg4; do
  PCS_DWT_FUSE=2 timeout 300 python scripts/dwt_time.py $W 32 20 >> $OUT/dwt_time.jsonl 2>> $OUT/dwt_time.err
done
cat $OUT/dwt_time.jsonl; cat $OUT/rc.txt
python scripts/dwt_time.py cfg2 32 3 > $OUT/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dwt_forward_l12 -s 2 -c 1 -o $OUT/prof_l12 python scripts/dwt_time.py cfg2 32 3 > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
