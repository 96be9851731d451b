#!/usr/bin/env python
"""Developer helper (GPU box): host-buffer (pinned) encode/decode times of the bench workload."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-image-and-video-codec_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import bench
from picsong_b200 import codec
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg = bench.WORKLOADS[name]
W, H, wl, lossy, qs, lutname, B, desc = cfg['W'], cfg['H'], cfg['wl'], cfg['lossy'], cfg['qs'], cfg['lut'], cfg['batch'], cfg['desc']
if len(sys.argv) > 2: B = int(sys.argv[2])
N = int(sys.argv[3]) if len(sys.argv) > 3 else B      # frames per call (may exceed the context's max_batch B)
lroot = bench.materialize_luts()
cd = codec.Codec(codec.Params(W, H, wl, bool(lossy), qs, max_batch=B), codec.load_lut(f"{lroot}/{lutname}/", wl))
import torch
fr = bench.gen_frames(cfg, range(N), torch.device('cuda')).cpu().numpy()
pin_px = torch.from_numpy(fr).pin_memory(); pin_out = torch.empty_like(pin_px).pin_memory()
pin_stream = torch.empty(N * cd.max_words * 3 // 5, dtype=torch.int16).pin_memory()
hp, ho, hs = pin_px.numpy(), pin_out.numpy(), pin_stream.numpy().view(np.uint16)
te, td = [], []
for i in range(8):
    t0 = time.perf_counter(); s, lens = cd.encode_frames(hp, 0, hs); t1 = time.perf_counter(); cd.decode_frames(s, lens, ho); t2 = time.perf_counter()
    if i >= 3: te.append(t1 - t0); td.append(t2 - t1)
print(json.dumps({"B": B, "N": N, "env": {k: v for k, v in os.environ.items() if k.startswith("PCS_")}, "enc_ms": 1e3 * float(np.median(te)), "dec_ms": 1e3 * float(np.median(td)), "e2e_mpix_s": W * H * N / 1e6 / float(np.median(te) + np.median(td)),
                  "lag": os.environ.get("PCS_PIPE_LAG"), "stream_MB": s.nbytes / 1e6}))
