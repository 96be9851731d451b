#!/usr/bin/env python
"""Developer helper (GPU box): per-stage device times of the hot path for one workload, with a
round-trip check.  PCS_LIB_PATH selects an alternative build of libpicsong_b200.so for A/B tests.

  python scripts/stage_time.py [cfg2|cfg3|cfg4|cfg1] [batch] [iters]
"""
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
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
lroot = bench.materialize_luts()
cd = codec.Codec(codec.Params(W, H, wl, bool(lossy), qs, max_batch=B), codec.load_lut(f"{lroot}/{lutname}/", wl))
px = bench.gen_frames(cfg, range(B), torch.device('cuda'))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn):
    ts = []
    for i in range(iters + 2):
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    return r, float(np.median(ts))


res = {"lib": os.environ.get("PCS_LIB_PATH", "default"), "workload": name, "batch": B, "mpix": W * H * B / 1e6}
coeff, res["dwt_fwd_ms"] = timed(lambda: cd.dwt_forward(px, trunc_to_int=True))
(cw, sizes), res["bpc_enc_ms"] = timed(lambda: cd.bpc_encode(coeff))
(stream, off), res["bsb_pack_ms"] = timed(lambda: cd.bsb_pack(cw, sizes, 0))
(cw2, sizes2), res["bsb_unpack_ms"] = timed(lambda: cd.bsb_unpack(stream, off))
q, res["bpc_dec_ms"] = timed(lambda: cd.bpc_decode(cw2, sizes2))
out, res["dwt_inv_ms"] = timed(lambda: cd.dwt_inverse(q))
res["coeff_roundtrip"] = bool(torch.equal(q, coeff))
res["pixels_equal"] = bool(torch.equal(out, px))
res["stream_words"] = int(off[-1].item())
res["stream_sha"] = __import__("hashlib").sha256(stream[: int(off[-1].item())].cpu().numpy().tobytes()).hexdigest()[:16]
A = cd.A
res["dwt_fwd_frac"] = 5.0 * A * B / (res["dwt_fwd_ms"] * 1e-3) / 1e9 / 6545.9
res["dwt_inv_frac"] = 5.0 * A * B / (res["dwt_inv_ms"] * 1e-3) / 1e9 / 6545.9
print(json.dumps(res))
