#!/usr/bin/env bash
# Round 2, GPU call R (1 GPU): explicit ncu pipe metrics of the BPC kernels on stage_time.py (ran plain in call P), launch list of the bench
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2r; mkdir -p $OUT; rm -f $OUT/*
timeout 300 python scripts/stage_time.py cfg4 32 2 > $OUT/plain.json 2>$OUT/err.txt; echo "plain rc=$?"
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,launch__registers_per_thread,smsp__inst_executed_op_branch.sum,l1tex__t_sector_hit_rate.pct
timeout 600 ncu --clock-control none -k regex:bpc_ -s 2 -c 4 --metrics $M --csv --log-file $OUT/bpc_metrics.csv python scripts/stage_time.py cfg4 32 2 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
tail -2 $OUT/ncu.log; wc -l $OUT/bpc_metrics.csv
# launch list of the bench command (ran plain in call Q), library kernels only (the synthetic-frame generator alone is >4000 torch launches)
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'dwt_|bpc_|bsb_|color_' -c 2500 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-continuity > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
wc -l $OUT/launches.csv; tail -2 $OUT/ncu_launches.log
