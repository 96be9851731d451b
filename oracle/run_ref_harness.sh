#!/usr/bin/env bash
# Test infrastructure; runs ON THE GPU BOX.  Stage dumps + stage-level timings of the reference.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/ref_harness; mkdir -p $OUT /tmp/rh
H=oracle/_ref/ref_harness; LUT=oracle/_ref/LUT
gen() { python - "$@" <<'PY'
import sys; sys.path.insert(0, "cuda-image-and-video-codec_b200")
from picsong_b200 import synth
kind, W, H, F, path = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
with open(path, "wb") as f:
    for i in range(max(F, 1)):
        a = synth.synth_frame(W, H, i if F else 0) if kind == "synth" else synth.noise_frame(W, H)
        f.write(a.tobytes())
PY
}
gen synth 256 256 3 /tmp/rh/v256.raw
$H 256 256 3 1 1.0 $LUT/video_lossy/ /tmp/rh/v256.raw 3 1 $OUT/v256_f3_wl3_q1 $OUT/v256_f3_wl3_q1.json
gen synth 512 512 0 /tmp/rh/s512.raw
$H 512 512 4 1 0.1 $LUT/n1_lossy/ /tmp/rh/s512.raw 0 1 $OUT/s512_wl4_q01 $OUT/s512_wl4_q01.json
$H 512 512 4 0 1.0 $LUT/n1_lossless/ /tmp/rh/s512.raw 0 1 $OUT/s512_wl4_ll $OUT/s512_wl4_ll.json
# timings (no dumps)
gen synth 4096 4096 0 /tmp/rh/s4096.raw
$H 4096 4096 5 1 1.0 $LUT/n1_lossy/ /tmp/rh/s4096.raw 0 6 - $OUT/t_cfg2_4096_wl5_q1.json
$H 4096 4096 5 0 1.0 $LUT/n1_lossless/ /tmp/rh/s4096.raw 0 6 - $OUT/t_cfg2_4096_wl5_ll.json
gen synth 7680 4320 0 /tmp/rh/s8k.raw
$H 7680 4320 6 0 1.0 $LUT/n1_lossless/ /tmp/rh/s8k.raw 0 4 - $OUT/t_cfg3_8k_wl6_ll.json
gen synth 3840 2160 8 /tmp/rh/v4k.raw
$H 3840 2160 5 1 1.0 $LUT/video_lossy/ /tmp/rh/v4k.raw 8 3 - $OUT/t_cfg4_4k_f8_wl5_q1.json
$H 512 512 4 0 1.0 $LUT/n1_lossless/ /tmp/rh/s512.raw 0 6 - $OUT/t_cfg1_512_wl4_ll.json
