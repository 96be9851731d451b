#!/usr/bin/env python
"""GPU box: wall-clock of the two command-line programs (ours and the reference's) on a raw 4K grey video held
in /dev/shm -- the whole application incl. file I/O, LUT parsing and allocation.  SURVEY 8f.3 (host pipeline)."""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-image-and-video-codec_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import lut_fixture
from picsong_b200 import synth

W, H, F = 3840, 2160, int(sys.argv[1]) if len(sys.argv) > 1 else 128
tmp = tempfile.mkdtemp(prefix="pcs_cli_", dir="/dev/shm")
lroot = tempfile.mkdtemp(prefix="pcs_lut_"); lut_fixture.materialize_all(lroot)
lut = f"{lroot}/video_lossy/"
raw = os.path.join(tmp, "v.raw")
distinct = [synth.synth_frame(W, H, f) for f in range(8)]
with open(raw, "wb") as f:
    for i in range(F):
        f.write(distinct[i % 8].tobytes())
res = {"frames": F, "mpix": W * H * F / 1e6}
for name, exe, ns in (("ours", os.path.join(ROOT, "cuda-image-and-video-codec_b200", "PICSONG_b200"), 2),
                      ("reference", os.path.join(ROOT, "oracle", "_ref", "PICSONG_ref"), 8)):
    enc, dec = os.path.join(tmp, name + ".pcs"), os.path.join(tmp, name + ".out")
    t0 = time.perf_counter()
    pe = subprocess.run([exe, "-cd", "0", "-wl", "5", "-cp", "2", "-type", "1", "-qs", "1.0", "-i", raw, "-o", enc, "-cbWidth", "64", "-cbHeight", "18",
                         "-xSize", str(W), "-ySize", str(H), "-video", "1", "-frames", str(F), "-numberOfStreams", str(ns), "-LUTFolder", lut, "-k", "0"],
                        capture_output=True, text=True)
    t1 = time.perf_counter()
    pd = subprocess.run([exe, "-cd", "1", "-i", enc, "-o", dec, "-LUTFolder", lut, "-video", "1", "-numberOfStreams", str(ns)], capture_output=True, text=True)
    t2 = time.perf_counter()
    res[name] = {"rc": [pe.returncode, pd.returncode], "enc_s": t1 - t0, "dec_s": t2 - t1, "enc_mpix_s": res["mpix"] / (t1 - t0), "dec_mpix_s": res["mpix"] / (t2 - t1),
                 "enc_bytes": os.path.getsize(enc) if os.path.exists(enc) else 0, "dec_bytes": os.path.getsize(dec) if os.path.exists(dec) else 0}
    if name == "reference":
        res["streams_identical"] = open(enc, "rb").read() == open(os.path.join(tmp, "ours.pcs"), "rb").read() if os.path.exists(enc) else None
print(json.dumps(res))
