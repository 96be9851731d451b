"""Developer helper (GPU box): prints where the CUDA 9/7 forward differs from the oracle."""
import sys, os, collections
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-image-and-video-codec_b200"), os.path.join(ROOT, "tests")]
from oracle import oracle as O
from picsong_b200 import codec, synth
import lut_fixture, tempfile
root = tempfile.mkdtemp(); lut_fixture.materialize_all(root)
W, H, wl, qs = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])) if len(sys.argv) > 4 else (256, 256, 3, 1.0)
plut = codec.load_lut(f"{root}/n1_lossy/", wl)
img = synth.synth_frame(W, H)
cd = codec.Codec(codec.Params(W, H, wl, True, qs), plut)
px = torch.from_numpy(img).cuda()[None]
g = cd.dwt_forward(px)[0].cpu().numpy()
o = O.dwt97_forward(O.pad_mirror(img), wl, qs)
d = np.argwhere(g.view(np.uint32) != o.view(np.uint32))
print("diffs", len(d), "max abs", np.abs(g - o).max())
for l in range(1, wl + 1):
    for sb, nm in ((1, "HL"), (2, "LH"), (3, "HH")) + (((0, "LL"),) if l == wl else ()):
        sx, sy = g.shape[1] >> l, g.shape[0] >> l
        ys = slice(sy, 2 * sy) if sb in (2, 3) else slice(0, sy)
        xs = slice(sx, 2 * sx) if sb in (1, 3) else slice(0, sx)
        a = g[ys, xs].view(np.uint32); b = o[ys, xs].view(np.uint32)
        dd = np.argwhere(a != b)
        print("level", l, nm, "diffs", len(dd), "rows", sorted(collections.Counter(dd[:, 0].tolist()).items())[:10],
              "cols", sorted(collections.Counter(dd[:, 1].tolist()).items())[:10], "maxabs", float(np.abs(g[ys, xs] - o[ys, xs]).max()))
