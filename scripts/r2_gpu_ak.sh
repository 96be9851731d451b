#!/usr/bin/env bash
# last check of round 2: smoke() + whole GPU suite on the clean rebuild of HEAD (1 GPU)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2ak; mkdir -p $OUT; rm -f $OUT/*
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -c 200 $OUT/pytest_gpu.log; tail -1 $OUT/smoke.log
