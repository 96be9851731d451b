#!/usr/bin/env bash
# Round 2, GPU call Q (1 GPU): whole GPU suite, default bench, then (after the plain runs exited 0) the ncu launch list of the
# bench command and explicit pipe metrics of the BPC kernels
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2q; mkdir -p $OUT; rm -f $OUT/*
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/rc.txt
tail -c 300 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_cfg4_n1.json 2> $OUT/bench_cfg4_n1.err; echo "bench rc=$?" >> $OUT/rc.txt
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-continuity > $OUT/bench_short.json 2> $OUT/bench_short.err; rc=$?; echo "bench short rc=$rc" >> $OUT/rc.txt
if [ $rc -eq 0 ]; then
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-continuity > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> $OUT/rc.txt
fi
cat $OUT/rc.txt; head -c 1500 $OUT/bench_cfg4_n1.json; tail -3 $OUT/bench_cfg4_n1.err; wc -l $OUT/launches.csv
