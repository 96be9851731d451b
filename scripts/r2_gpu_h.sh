#!/usr/bin/env bash
# Round 2, GPU call H (1 GPU): reference behaviour for -bps / -signedOrUnsigned (f4), CLI tests + whole-application wall
# clock against the reference executable, BPC occupancy A/B, ncu local-memory / issue metrics of the coder kernels.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2h; mkdir -p $OUT; rm -f $OUT/*
bash scripts/r2_ref_f4.sh > $OUT/f4.log 2>&1
timeout 600 python -m pytest tests/test_gpu_cli.py -m gpu -x -q > $OUT/pytest_cli.log 2>&1; echo "pytest cli rc=$?" >> $OUT/rc.txt
tail -c 500 $OUT/pytest_cli.log
timeout 900 python scripts/cli_video_bench.py 64 600 > $OUT/cli_video.json 2> $OUT/cli_video.err; echo "cli bench rc=$?" >> $OUT/rc.txt
timeout 300 python scripts/stage_time.py cfg4 32 10 >> $OUT/stage_time.jsonl 2>> $OUT/stage_time.err
for V in build/variants/lib_*.so; do [ -f "$V" ] && PCS_LIB_PATH=$PWD/$V timeout 300 python scripts/stage_time.py cfg4 32 10 >> $OUT/stage_time.jsonl 2>> $OUT/stage_time.err; done
cut -c1-420 $OUT/stage_time.jsonl
python scripts/stage_time.py cfg4 32 2 > $OUT/plain.log 2>&1 && ncu --clock-control none -k regex:bpc_ -c 6 --csv --log-file $OUT/bpc_metrics.csv \
  --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
  python scripts/stage_time.py cfg4 32 2 > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log; cat $OUT/rc.txt; head -c 2500 $OUT/cli_video.json; tail -3 $OUT/cli_video.err
