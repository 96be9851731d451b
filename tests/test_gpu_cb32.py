"""GPU parity of 64 x 32 codeblocks (PCS_F_CB_LENGTH_32; BASELINE configs[2]'s sweep point = the reference rebuilt with
-DCBLOCK_LENGTH=32): the CUDA path against the CPU oracle stage by stage, and against streams of that reference build
(tests/golden/cb32_ref.npz) where present."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import oracle as O  # noqa: E402  (checker only)
from picsong_b200 import codec, synth  # noqa: E402

# name, kind, W, H, wl, lossy, qs
CASES = [
    ("cb32_s256x160_wl3_ll", "synth", 256, 160, 3, 0, 1.0),     # H = 5 x 32: no padding with 32-row blocks, 192 with 64-row blocks
    ("cb32_s512x288_wl3_q1", "synth", 512, 288, 3, 1, 1.0),
    ("cb32_pad200x150_wl2_ll", "synth", 200, 150, 2, 0, 1.0),  # padded to 256 x 160
    ("cb32_noise128x96_wl2_ll", "noise", 128, 96, 2, 0, 1.0),  # blocks that need the raw dump (size 2048)
    ("cb32_s1024x544_wl4_q05", "synth", 1024, 544, 4, 1, 0.5),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cb32_against_oracle(case, lut_root):
    name, kind, W, H, wl, lossy, qs = case
    folder = "n1_lossy" if lossy else "n1_lossless"
    olut, plut = O.lut_load(f"{lut_root}/{folder}/", wl), codec.load_lut(f"{lut_root}/{folder}/", wl)
    img = synth.synth_frame(W, H) if kind == "synth" else synth.noise_frame(W, H)
    cd = codec.Codec(codec.Params(W, H, wl, bool(lossy), qs, flags=codec.PCS_F_CB_LENGTH_32), plut)
    try:
        O.set_cb_rows(32)
        Wa, Ha = O.adapted_dims(W, H)
        assert (cd.Wa, cd.Ha) == (Wa, Ha) and cd.ncb == (Wa // 64) * (Ha // 32)
        px = torch.from_numpy(img).cuda()[None]
        opad = O.pad_mirror(img)
        oq = O.truncate(O.dwt97_forward(opad, wl, qs)) if lossy else O.dwt53_forward(opad, wl)
        coeff = cd.dwt_forward(px, trunc_to_int=True)
        assert np.array_equal(coeff[0].cpu().numpy(), oq), "coefficients differ"
        ocs, osz = O.bpc_encode(oq, wl, olut)
        cw, sizes = cd.bpc_encode(coeff)
        assert np.array_equal(sizes[0].cpu().numpy(), osz), "codeblock sizes differ"
        if (osz < 2047).all():  # the oracle follows the reference, which is only defined while no block passes 2046 codewords
            assert np.array_equal(cd.export_ref_codestream(cw, sizes)[0].cpu().numpy(), ocs), "sparse codestream differs"
        stream, off = cd.bsb_pack(cw, sizes, 0)
        n = int(off.cpu().numpy()[1])
        if (osz < 2047).all():
            assert np.array_equal(stream[:n].cpu().numpy().view(np.uint16), O.bsb_pack(ocs, osz, O.header_pack(W, H, wl, lossy, qs, 0)))
        cw_b, sizes_b = cd.bsb_unpack(stream, off)
        q = cd.bpc_decode(cw_b, sizes_b)
        assert np.array_equal(q[0].cpu().numpy(), oq), "decoded coefficients differ"
        out = cd.dwt_inverse(q)[0].cpu().numpy()
        if not lossy:
            assert np.array_equal(out, img)
        # host-buffer calls
        s, lens = cd.encode_frames(img[None])
        assert np.array_equal(s, stream[:n].cpu().numpy().view(np.uint16))
        assert np.array_equal(cd.decode_frames(s, lens)[0], out)
        gold = os.path.join(os.path.dirname(__file__), "golden", "cb32_ref.npz")
        if os.path.exists(gold):
            g = np.load(gold)
            if name + ".stream" in g:
                assert np.array_equal(s, g[name + ".stream"]), f"{name}: stream differs from the reference's 64x32 build"
    finally:
        O.set_cb_rows(64)
        cd.close()


def test_cb32_full_size_config3_matches_reference_build(lut_root):
    """BASELINE configs[2] at its 64 x 32 sweep point: 7680x4320 lossless, wl = 5 (the deepest the reference's own 64x32
    build round-trips at: its wl = 6 run does not reproduce the input, tests/golden/ref_manifest.json).  sha256 of the
    reference build's stream, and the identity round trip."""
    import hashlib
    import json
    e = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_manifest.json")))["cfg3_8k_wl5_ll_cb32"]
    plut = codec.load_lut(f"{lut_root}/{e['lut']}/", e["wl"])
    cd = codec.Codec(codec.Params(e["W"], e["H"], e["wl"], False, 1.0, flags=codec.PCS_F_CB_LENGTH_32), plut)
    assert (cd.Wa, cd.Ha) == (7680, 4320)          # no padding rows with 32-row blocks (4352 with 64-row blocks)
    img = synth.synth_frame(e["W"], e["H"])
    stream, lens = cd.encode_frames(img[None])
    assert stream.size * 2 == e["enc_len"]
    assert hashlib.sha256(stream.tobytes()).hexdigest() == e["enc_sha256"]
    assert np.array_equal(cd.decode_frames(stream, lens)[0], img)
    cd.close()
