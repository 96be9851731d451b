#!/usr/bin/env bash
# host-buffer pipeline A/B on the 600-frame call (1 GPU) + CLI wall clock with the new start-up detail
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2k; mkdir -p $OUT; rm -f $OUT/*
run() { env "$@" timeout 300 python scripts/e2e_time.py cfg4 32 600 >> $OUT/e2e.jsonl 2>> $OUT/e2e.err; }
run X=1
run PCS_PIPE_CHUNK=8
run PCS_PIPE_CHUNK=8 PCS_PIPE_DEPTH=3
run PCS_PIPE_CHUNK=8 PCS_PIPE_DEPTH=4
run PCS_PIPE_CHUNK=4 PCS_PIPE_DEPTH=3
run PCS_PIPE_CHUNK=6 PCS_PIPE_DEPTH=2
run PCS_PIPE_CHUNK=8 PCS_PIPE_LAG=3
cut -c1-300 $OUT/e2e.jsonl; tail -3 $OUT/e2e.err
CLI_BENCH_REPS=3 timeout 900 python scripts/cli_video_bench.py 64 > $OUT/cli_video64.json 2> $OUT/cli_video64.err; echo "cli rc=$?"
python - <<'PY'
import json
r = json.load(open("gpurun_out/r2k/cli_video64.json"))["64"]
print("ours", r["ours"]["all_walls_s"], r["ours"]["enc_own"], r["ours"]["dec_own"])
for x in r["reference_all"]: print("ref", x["streams"], x["all_walls_s"], x["enc_own"], x["dec_own"])
PY
