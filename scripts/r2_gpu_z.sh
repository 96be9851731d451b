#!/usr/bin/env bash
# k-mode coder with empty-call-site skip: parity (k-factor + CLI k test), bench --k 1 (1 GPU)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2z; mkdir -p $OUT; rm -f $OUT/*
timeout 1500 python -m pytest tests/test_gpu_kfactor.py tests/test_gpu_cli.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -c 300 $OUT/pytest.log
timeout 600 python bench.py --workload cfg4 --k 1.0 --steps 3 --warmup 3 --no-cpu-baseline --no-continuity > $OUT/bench_cfg4_k1.json 2> $OUT/bench_cfg4_k1.err; echo "k1 rc=$?"
python -c "
import json; d=json.load(open('$OUT/bench_cfg4_k1.json')); print(d['value'], d['e2e']['value'], d['stage_ms'])"
