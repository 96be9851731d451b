#!/usr/bin/env bash
# final build of round 2 (1 GPU): GPU suite, default bench
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2aj; mkdir -p $OUT; rm -f $OUT/*
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/rc.txt
tail -c 200 $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?" >> $OUT/rc.txt
cat $OUT/rc.txt; head -c 2500 $OUT/bench_default.json; tail -3 $OUT/bench_default.err
