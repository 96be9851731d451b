#!/usr/bin/env bash
# fused DWT: decoupled hand-over (mbarrier on the heavy warps' side) + L2 hints: parity, then timing A/B (1 GPU)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2n; mkdir -p $OUT; rm -f $OUT/*
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_levels or golden or cfg2" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -c 600 $OUT/pytest.log
for w in cfg2 cfg4; do
  PCS_DWT_FUSE_INV=0 timeout 300 python scripts/dwt_time.py $w > $OUT/dwt_${w}_perlevel.json 2>$OUT/err.txt; echo "perlevel-inv rc=$?"; cat $OUT/dwt_${w}_perlevel.json
  timeout 300 python scripts/dwt_time.py $w > $OUT/dwt_${w}_fused.json 2>>$OUT/err.txt; echo "default rc=$?"; cat $OUT/dwt_${w}_fused.json
  for v in nodecouple nohint hint6; do
    PCS_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 300 python scripts/dwt_time.py $w > $OUT/dwt_${w}_$v.json 2>>$OUT/err.txt; echo "$v rc=$?"; cat $OUT/dwt_${w}_$v.json
  done
done
tail -5 $OUT/err.txt
