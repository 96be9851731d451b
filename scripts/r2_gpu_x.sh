#!/usr/bin/env bash
# Round 2, GPU call X (1 GPU): whole GPU suite, default bench, CLI wall clock; then (after the plain runs exited 0) BPC pipe metrics
# and the launch list of the bench command under ncu
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2x; mkdir -p $OUT; rm -f $OUT/*
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/rc.txt
tail -c 300 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_cfg4_n1.json 2> $OUT/bench_cfg4_n1.err; echo "bench rc=$?" >> $OUT/rc.txt
timeout 600 python bench.py --workload cfg4 --k 1.0 --steps 3 --warmup 3 --no-cpu-baseline --no-continuity > $OUT/bench_cfg4_k1.json 2> $OUT/bench_cfg4_k1.err; echo "k1 rc=$?" >> $OUT/rc.txt
timeout 300 python scripts/stage_time.py cfg4 32 > $OUT/stage_cfg4.json 2>$OUT/err.txt; echo "stage rc=$?" >> $OUT/rc.txt
timeout 300 python scripts/stage_time.py cfg2 32 > $OUT/stage_cfg2.json 2>>$OUT/err.txt
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,launch__registers_per_thread,smsp__inst_executed_op_branch.sum,l1tex__t_sector_hit_rate.pct
timeout 600 ncu --clock-control none -k regex:bpc_ -s 2 -c 4 --metrics $M --csv --log-file $OUT/bpc_metrics.csv python scripts/stage_time.py cfg4 32 2 > $OUT/ncu.log 2>&1; echo "ncu bpc rc=$?" >> $OUT/rc.txt
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'dwt_|bpc_|bsb_|color_' -c 2500 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-continuity > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> $OUT/rc.txt
cat $OUT/rc.txt; head -c 1200 $OUT/bench_cfg4_n1.json; echo; cat $OUT/stage_cfg4.json
