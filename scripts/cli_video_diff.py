#!/usr/bin/env python
"""GPU box, developer helper: where do two CLI encodes of the same raw video differ (frame by frame)?"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-image-and-video-codec_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import lut_fixture
from picsong_b200 import synth
W, H, F = 3840, 2160, int(sys.argv[1]) if len(sys.argv) > 1 else 64
tmp = tempfile.mkdtemp(prefix="pcs_cli_", dir="/dev/shm")
lroot = tempfile.mkdtemp(prefix="pcs_lut_"); lut_fixture.materialize_all(lroot)
lut = f"{lroot}/video_lossy/"
raw = os.path.join(tmp, "v.raw")
distinct = [synth.synth_frame(W, H, f) for f in range(8)]
with open(raw, "wb") as f:
    for i in range(F):
        f.write(distinct[i % 8].tobytes())
outs = {}
for name, exe, ns in (("ours", os.path.join(ROOT, "cuda-image-and-video-codec_b200", "PICSONG_b200"), 2),
                      ("ref8", os.path.join(ROOT, "oracle", "_ref", "PICSONG_ref"), 8),
                      ("ref2", os.path.join(ROOT, "oracle", "_ref", "PICSONG_ref"), 2)):
    enc = os.path.join(tmp, name + ".pcs")
    p = subprocess.run([exe, "-cd", "0", "-wl", "5", "-cp", "2", "-type", "1", "-qs", "1.0", "-i", raw, "-o", enc, "-cbWidth", "64", "-cbHeight", "18",
                        "-xSize", str(W), "-ySize", str(H), "-video", "1", "-frames", str(F), "-numberOfStreams", str(ns), "-LUTFolder", lut, "-k", "0"],
                       capture_output=True, text=True)
    sizes = [int(t) for t in open(enc + "_SIZE").read().split(",")]
    outs[name] = (np.fromfile(enc, dtype=np.uint16), sizes)
    print(name, p.returncode, len(sizes), sum(sizes), outs[name][0].size)
a, sa = outs["ours"]
for other in ("ref8", "ref2"):
    b, sb = outs[other]
    bad = [i for i, (x, y) in enumerate(zip(sa, sb)) if x != y]
    print(other, "frames with different size:", bad[:10], [(sa[i], sb[i]) for i in bad[:5]])
    oa = np.concatenate([[0], np.cumsum(sa)]); ob = np.concatenate([[0], np.cumsum(sb)])
    for i in range(min(len(sa), len(sb))):
        fa, fb = a[oa[i]:oa[i + 1]], b[ob[i]:ob[i + 1]]
        if fa.size != fb.size or not np.array_equal(fa, fb):
            n = min(fa.size, fb.size); d = np.nonzero(fa[:n] != fb[:n])[0]
            print("  frame", i, "sizes", fa.size, fb.size, "first diff word", d[:5], fa[d[:5]], fb[d[:5]])
            break
    else:
        print("  all frames identical")
