#!/usr/bin/env python
"""Developer helper (GPU box): device time of the forward / inverse DWT alone for one workload (L2 flushed between
iterations), for A/B runs of the PCS_DWT* environment switches.

  python scripts/dwt_time.py [cfg2|cfg3|cfg4|cfg1] [batch] [iters]
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-image-and-video-codec_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch

import bench
from picsong_b200 import codec

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg = bench.WORKLOADS[name]
W, H, wl, lossy, qs, lutname, B, desc = cfg['W'], cfg['H'], cfg['wl'], cfg['lossy'], cfg['qs'], cfg['lut'], cfg['batch'], cfg['desc']
if len(sys.argv) > 2:
    B = int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
lroot = bench.materialize_luts()
cd = codec.Codec(codec.Params(W, H, wl, bool(lossy), qs, max_batch=B), codec.load_lut(f"{lroot}/{lutname}/", wl))
px = bench.gen_frames(cfg, range(B), torch.device('cuda'))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn):
    ts = []
    for i in range(iters + 3):
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    return r, float(np.median(ts)), float(np.min(ts))


coeff, fwd, fwd_min = timed(lambda: cd.dwt_forward(px, trunc_to_int=True))
out, inv, inv_min = timed(lambda: cd.dwt_inverse(coeff))
env = {k: v for k, v in os.environ.items() if k.startswith("PCS_")}
A = cd.A
print(json.dumps({"workload": name, "batch": B, "env": env, "fwd_ms": round(fwd, 4), "fwd_min_ms": round(fwd_min, 4), "inv_ms": round(inv, 4),
                  "fwd_frac": round(5.0 * A * B / (fwd * 1e-3) / 1e9 / 6545.9, 4), "inv_frac": round(5.0 * A * B / (inv * 1e-3) / 1e9 / 6545.9, 4),
                  "coeff_sha": hashlib.sha256(coeff.cpu().numpy().tobytes()).hexdigest()[:12],
                  "px_sha": hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:12]}))
