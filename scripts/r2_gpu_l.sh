#!/usr/bin/env bash
# CLI tests + whole-application wall clock (64 and 600 frames) with pageable buffers for short jobs (1 GPU)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2l; mkdir -p $OUT; rm -f $OUT/*
timeout 600 python -m pytest tests/test_gpu_cli.py tests/test_gpu_bitdepth.py -m gpu -x -q > $OUT/pytest_cli.log 2>&1; echo "pytest rc=$?"
tail -c 300 $OUT/pytest_cli.log
CLI_BENCH_REPS=3 timeout 1200 python scripts/cli_video_bench.py 64 600 > $OUT/cli_video.json 2> $OUT/cli_video.err; echo "cli rc=$?"
python - <<'PY'
import json
r = json.load(open("gpurun_out/r2l/cli_video.json"))
for F, v in r.items():
    print(F, "ours", v["ours"]["all_walls_s"], v["ours"]["enc_own"], v["ours"]["dec_own"])
    for x in v["reference_all"]: print(F, "ref", x["streams"], x["all_walls_s"], x["enc_own"].get("no_alloc_s"), x["dec_own"].get("no_alloc_s"))
PY
