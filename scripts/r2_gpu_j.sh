#!/usr/bin/env bash
# Round 2, GPU call J (1 GPU): whole GPU suite, configs[2] sweep (ours, 64x32), complexity-scalable mode on both arms
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2j; mkdir -p $OUT; rm -f $OUT/*
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/rc.txt
tail -c 600 $OUT/pytest_gpu.log
for WL in 5; do for CBL in 64 32; do
  timeout 600 python bench.py --workload cfg3 --cb-length $CBL --wl $WL --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg3_cb${CBL}_wl$WL.json 2> $OUT/bench_cfg3_cb${CBL}_wl$WL.err; echo "cfg3 cb$CBL rc=$?" >> $OUT/rc.txt
done; done
timeout 900 python bench.py --workload cfg4 --k 1.0 --steps 3 --warmup 3 --no-cpu-baseline --no-continuity > $OUT/bench_cfg4_k1.json 2> $OUT/bench_cfg4_k1.err; echo "k1 ours rc=$?" >> $OUT/rc.txt
timeout 900 python bench.py --impl reference --workload cfg4 --k 1.0 --steps 3 --warmup 3 > $OUT/bench_ref_cfg4_k1.json 2> $OUT/bench_ref_cfg4_k1.err; echo "k1 ref rc=$?" >> $OUT/rc.txt
cat $OUT/rc.txt
for f in $OUT/bench_*.json; do echo $f; head -c 700 $f; echo; done; tail -3 $OUT/bench_cfg4_k1.err
