#!/usr/bin/env bash
# forward DWT A/B: default library and the variant libraries under build/variants (1 GPU); parity first
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2f; mkdir -p $OUT; rm -f $OUT/*
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_levels" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" > $OUT/rc.txt
tail -c 300 $OUT/pytest.log
for W in cfg2 cfg4; do
  timeout 300 python scripts/dwt_time.py $W 32 20 >> $OUT/dwt_time.jsonl 2>> $OUT/dwt_time.err
  for V in build/variants/lib_*.so; do
    [ -f "$V" ] && PCS_LIB_PATH=$PWD/$V timeout 300 python scripts/dwt_time.py $W 32 20 >> $OUT/dwt_time.jsonl 2>> $OUT/dwt_time.err
  done
done
cut -c1-260 $OUT/dwt_time.jsonl; cat $OUT/rc.txt
if [ "${1:-}" = "ncu" ]; then
python scripts/dwt_time.py cfg2 32 3 > $OUT/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dwt_forward_l12 -s 2 -c 1 -o $OUT/prof_l12 python scripts/dwt_time.py cfg2 32 3 > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
fi
