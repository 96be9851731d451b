#!/usr/bin/env bash
# Round 2, GPU call AH (N GPUs): multi-rank tests + sharded bench at 96 frames per launch
set -u
N=${1:-2}
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2ah_n$N; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; nproc >> $OUT/topo.txt; free -g >> $OUT/topo.txt
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > $OUT/pytest_multirank.log 2>&1; echo "pytest rc=$?" >> $OUT/rc.txt
for WLD in cfg4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus $N --steps 5 --warmup 3 --workload $WLD > $OUT/bench_${WLD}_n$N.json 2> $OUT/bench_${WLD}_n$N.err; echo "bench $WLD rc=$?" >> $OUT/rc.txt
done
tail -c 400 $OUT/pytest_multirank.log; cat $OUT/rc.txt; head -c 1800 $OUT/bench_cfg4_n$N.json; tail -5 $OUT/bench_cfg4_n$N.err
