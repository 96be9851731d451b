#!/usr/bin/env bash
# Round 2, GPU call I (1 GPU): 64x32 codeblocks.  Reference (-DCBLOCK_LENGTH=32 build) runs that pin the small cases and
# probe configs[2] (7680x4320, wl 6 and 5), our GPU tests for the new paths, both arms of the configs[2] sweep.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2i; mkdir -p $OUT /tmp/r2i; rm -f $OUT/*
R=oracle/_ref/PICSONG_ref_cb32; LUT=oracle/_ref/LUT
python - <<'PY'
import sys; sys.path[:0] = ["cuda-image-and-video-codec_b200", "tests"]
from picsong_b200 import synth
import test_gpu_cb32 as t
for name, kind, W, H, wl, lossy, qs in t.CASES:
    (synth.synth_frame(W, H) if kind == "synth" else synth.noise_frame(W, H)).tofile(f"/tmp/r2i/{name}.raw")
synth.synth_frame(7680, 4320).tofile("/tmp/r2i/cfg3.raw")
PY
run_ref() {  # name W H wl type qs lut
  timeout 120 $R -cd 0 -wl $4 -cp 2 -type $5 -qs $6 -i /tmp/r2i/$1.raw -o $OUT/$1.pcs -cbWidth 64 -cbHeight 18 -xSize $2 -ySize $3 -video 0 -LUTFolder $LUT/$7/ -k 0 > $OUT/$1.enc.log 2>&1
  echo "$1 enc rc=$?" >> $OUT/rc.txt
  timeout 120 $R -cd 1 -i $OUT/$1.pcs -o $OUT/$1.dec.pgm -LUTFolder $LUT/$7/ -video 0 > $OUT/$1.dec.log 2>&1
  echo "$1 dec rc=$?" >> $OUT/rc.txt
}
run_ref cb32_s256x160_wl3_ll 256 160 3 0 1.0 n1_lossless
run_ref cb32_s512x288_wl3_q1 512 288 3 1 1.0 n1_lossy
run_ref cb32_pad200x150_wl2_ll 200 150 2 0 1.0 n1_lossless
run_ref cb32_noise128x96_wl2_ll 128 96 2 0 1.0 n1_lossless
run_ref cb32_s1024x544_wl4_q05 1024 544 4 1 0.5 n1_lossy
cp /tmp/r2i/cfg3.raw /tmp/r2i/cfg3_wl6.raw; cp /tmp/r2i/cfg3.raw /tmp/r2i/cfg3_wl5.raw
run_ref cfg3_wl6 7680 4320 6 0 1.0 n1_lossless
run_ref cfg3_wl5 7680 4320 5 0 1.0 n1_lossless
python - <<'PY' > gpurun_out/r2i/cfg3_check.txt 2>&1
import numpy as np, hashlib
inp = np.fromfile("/tmp/r2i/cfg3.raw", dtype=np.uint8)
for n in ("cfg3_wl6", "cfg3_wl5"):
    try:
        blob = open(f"gpurun_out/r2i/{n}.dec.pgm", "rb").read()
        body = np.frombuffer(blob.split(b"\n", 3)[3], dtype=np.uint8)
        enc = open(f"gpurun_out/r2i/{n}.pcs", "rb").read()
        print(n, "enc bytes", len(enc), "sha256", hashlib.sha256(enc).hexdigest(), "round trip:", body.size == inp.size and bool(np.array_equal(body, inp)),
              "differing px", int((body[:inp.size] != inp[:body.size]).sum()) if body.size else None)
    except Exception as e:
        print(n, "failed", e)
PY
rm -f $OUT/cfg3_wl*.pcs $OUT/cfg3_wl*.dec.pgm
cat $OUT/cfg3_check.txt; cat $OUT/rc.txt
timeout 900 python -m pytest tests/test_gpu_cb32.py tests/test_gpu_bitdepth.py tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/rc.txt
tail -c 1200 $OUT/pytest.log
for CBL in 64 32; do for WL in 6 5; do
  timeout 600 python bench.py --workload cfg3 --cb-length $CBL --wl $WL --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg3_cb${CBL}_wl$WL.json 2> $OUT/bench_cfg3_cb${CBL}_wl$WL.err
  timeout 600 python bench.py --impl reference --workload cfg3 --cb-length $CBL --wl $WL --steps 5 --warmup 3 > $OUT/bench_ref_cfg3_cb${CBL}_wl$WL.json 2> $OUT/bench_ref_cfg3_cb${CBL}_wl$WL.err
done; done
for f in $OUT/bench_*cfg3*.json; do echo $f; head -c 420 $f; echo; done; tail -3 $OUT/bench_cfg3_cb32_wl5.err
