#!/usr/bin/env bash
# Round 2, GPU call A (1 GPU): box facts, reference runs that pin the rare coder branches and the determinism of the
# reference on configs[1], the GPU test suite, the new default bench on both arms.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2a; mkdir -p $OUT /tmp/r2a
{ df -h /dev/shm /tmp; nproc; free -g; lscpu | head -25; nvidia-smi topo -m; ulimit -l; } > $OUT/box.txt 2>&1
# ---- reference coder / decoder on the synthetic coefficient planes (tests/coef_planes.py) ----
python - <<'PY'
import sys; sys.path[:0] = ["tests"]
import coef_planes
for n in coef_planes.CASES:
    coef_planes.plane(n).tofile(f"/tmp/r2a/{n}.i32")
open("/tmp/r2a/dummy.raw", "wb").write(bytes(coef_planes.W * coef_planes.H))
PY
mkdir -p $OUT/coef
for n in dense12 dense14 dense18 sparse17 sparse20 sparse22 mixed; do
  timeout 120 oracle/_ref/ref_harness 128 128 2 0 1.0 oracle/_ref/LUT/n1_lossless/ /tmp/r2a/dummy.raw 0 1 $OUT/coef/$n $OUT/coef/$n.json coef=/tmp/r2a/$n.i32 > $OUT/coef/$n.log 2>&1
  echo "coef $n rc=$?" >> $OUT/coef/rc.txt
done
# ---- is the reference deterministic on configs[1] lossy?  5 runs per qs, sha256 of its stream ----
timeout 900 python scripts/ref_determinism.py $OUT/ref_determinism.json > $OUT/ref_determinism.log 2>&1
# ---- GPU tests ----
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/rc.txt
# ---- bench, both arms ----
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_cfg4_n1.json 2> $OUT/bench_cfg4_n1.err; echo "bench rc=$?" >> $OUT/rc.txt
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_ref_cfg4.json 2> $OUT/bench_ref_cfg4.err; echo "ref rc=$?" >> $OUT/rc.txt
tail -c 600 $OUT/pytest_gpu.log; cat $OUT/rc.txt; head -c 1500 $OUT/bench_cfg4_n1.json; tail -5 $OUT/bench_cfg4_n1.err
