#!/usr/bin/env python
"""Test infrastructure, runs ON THE GPU BOX: is the reference executable deterministic on BASELINE configs[1]
(4096x4096 lossy 9/7)?  Runs oracle/_ref/PICSONG_ref several times per quantiser step on the same input and records
the sha256 of each stream, next to the sha256 of our stream.  Output: one JSON file (argv[1])."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-image-and-video-codec_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402

from picsong_b200 import codec, synth  # noqa: E402


def main(out_path, runs=5):
    W = H = 4096
    tmp = tempfile.mkdtemp(prefix="pcs_det_")
    raw = os.path.join(tmp, "in.raw")
    img = synth.synth_frame(W, H)
    img.tofile(raw)
    exe = os.path.join(ROOT, "oracle", "_ref", "PICSONG_ref")
    lutdir = os.path.join(ROOT, "oracle", "_ref", "LUT", "n1_lossy") + "/"
    res = {}
    for qs in (1.0, 0.5, 0.1):
        shas, lens = [], []
        for r in range(runs):
            enc = os.path.join(tmp, f"o_{qs}_{r}.pcs")
            subprocess.run([exe, "-cd", "0", "-wl", "5", "-cp", "2", "-type", "1", "-qs", repr(qs), "-i", raw, "-o", enc, "-cbWidth", "64",
                            "-cbHeight", "18", "-xSize", str(W), "-ySize", str(H), "-video", "0", "-LUTFolder", lutdir, "-numberOfStreams", "1",
                            "-k", "0"], capture_output=True, text=True)
            data = open(enc, "rb").read() if os.path.exists(enc) else b""
            shas.append(hashlib.sha256(data).hexdigest()); lens.append(len(data))
            if os.path.exists(enc):
                os.remove(enc)
        cd = codec.Codec(codec.Params(W, H, 5, True, qs), codec.load_lut(lutdir, 5))
        s, _ = cd.encode_frames(img[None])
        cd.close()
        res[str(qs)] = {"reference_sha256": shas, "reference_len": lens, "distinct": len(set(shas)),
                        "ours_sha256": hashlib.sha256(s.tobytes()).hexdigest(), "ours_len": int(s.size * 2),
                        "ours_equals_a_reference_run": hashlib.sha256(s.tobytes()).hexdigest() in shas}
    json.dump(res, open(out_path, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
