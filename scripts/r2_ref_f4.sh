#!/usr/bin/env bash
# Test infrastructure, runs ON THE GPU BOX: what does the unmodified reference do with -bps 16 and -signedOrUnsigned 1
# ?  Encodes + decodes a 256x192 synthetic still, lossless and lossy, records exit codes, streams and outputs.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2f4; mkdir -p $OUT /tmp/f4; rm -f $OUT/*
R=oracle/_ref/PICSONG_ref; LUT=oracle/_ref/LUT
python - <<'PY'
import sys; sys.path.insert(0, "cuda-image-and-video-codec_b200")
from picsong_b200 import synth
synth.synth_frame(256, 192).tofile("/tmp/f4/in.raw")
PY
for case in "bps16_ll -type 0 -bps 16 n1_lossless" "bps16_q1 -type 1 -bps 16 n1_lossy" "signed_ll -type 0 -signedOrUnsigned 1 n1_lossless" "signed_q1 -type 1 -signedOrUnsigned 1 n1_lossy" "bps12_ll -type 0 -bps 12 n1_lossless"; do
  set -- $case; name=$1; lut=${@: -1}; flags="${@:2:$#-2}"
  timeout 60 $R -cd 0 -wl 3 -cp 2 $flags -qs 1.0 -i /tmp/f4/in.raw -o $OUT/$name.pcs -cbWidth 64 -cbHeight 18 -xSize 256 -ySize 192 -video 0 -LUTFolder $LUT/$lut/ -k 0 > $OUT/$name.enc.log 2>&1
  echo "$name enc rc=$?" >> $OUT/rc.txt
  timeout 60 $R -cd 1 -i $OUT/$name.pcs -o $OUT/$name.dec.pgm -LUTFolder $LUT/$lut/ -video 0 > $OUT/$name.dec.log 2>&1
  echo "$name dec rc=$?" >> $OUT/rc.txt
done
cp /tmp/f4/in.raw $OUT/in.raw
cat $OUT/rc.txt; ls -la $OUT | head -30
