#!/usr/bin/env bash
# Round 2, GPU call G (1 GPU): whole GPU suite, default bench on both arms, ncu launch list + full capture of the DWT call
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2g; mkdir -p $OUT; rm -f $OUT/*
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/rc.txt
tail -c 400 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_cfg4_n1.json 2> $OUT/bench_cfg4_n1.err; echo "bench rc=$?" >> $OUT/rc.txt
python scripts/dwt_time.py cfg4 32 3 > $OUT/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dwt_forward -s 8 -c 4 -o $OUT/prof_dwt_fwd_cfg4 python scripts/dwt_time.py cfg4 32 3 > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log; cat $OUT/rc.txt; head -c 1200 $OUT/bench_cfg4_n1.json; tail -3 $OUT/bench_cfg4_n1.err
