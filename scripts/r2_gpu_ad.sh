#!/usr/bin/env bash
# Round 2, final check of the committed build (1 GPU): smoke(), whole GPU suite, default bench on both arms
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2ad; mkdir -p $OUT; rm -f $OUT/*
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/rc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/rc.txt
tail -c 300 $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?" >> $OUT/rc.txt
timeout 1500 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench ref rc=$?" >> $OUT/rc.txt
cat $OUT/rc.txt; tail -2 $OUT/smoke.log; head -c 600 $OUT/bench_default.json; echo; head -c 900 $OUT/bench_reference.json
