#!/usr/bin/env bash
# BPC: encoder at 5 CTAs/SM, decoder with pinned context words: parity, stage times, then ncu pipe / local-memory metrics
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2t; mkdir -p $OUT; rm -f $OUT/*
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_coder_branches.py tests/test_gpu_cb32.py tests/test_gpu_bitdepth.py tests/test_gpu_rgb.py tests/test_gpu_kfactor.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -c 300 $OUT/pytest.log
for w in cfg4 cfg2; do timeout 300 python scripts/stage_time.py $w 32 > $OUT/stage_${w}.json 2>$OUT/err.txt; echo "$w rc=$?"; cat $OUT/stage_${w}.json; done
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,launch__registers_per_thread,smsp__inst_executed_op_branch.sum,l1tex__t_sector_hit_rate.pct
timeout 600 ncu --clock-control none -k regex:bpc_ -s 2 -c 4 --metrics $M --csv --log-file $OUT/bpc_metrics.csv python scripts/stage_time.py cfg4 32 2 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
tail -2 $OUT/ncu.log | cut -c1-300
