"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C ABI, against
 (a) the CPU oracle on the same seeded inputs, stage by stage, and
 (b) the committed outputs of the unmodified reference binary (tests/golden).
Bit-exact for every integer/byte result (5/3 coefficients, quantisation indices, codewords, packed
stream, decoded pixels of lossless streams).  9/7: forward coefficients bit-exact against the oracle;
decoded pixels bit-exact against the oracle and within +-1 on <= 1e-4 of the pixels against the
reference (its inverse uses -use_fast_math divisions)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import oracle as O  # noqa: E402  (checker only)
from picsong_b200 import codec, synth  # noqa: E402

LUTS = {(0, 0): "n1_lossless", (1, 0): "n1_lossy", (0, 1): "video_lossless", (1, 1): "video_lossy"}

# name, kind, W, H, wl, lossy, qs
CASES = [
    ("s128_wl2_ll", "synth", 128, 128, 2, 0, 1.0),
    ("s256x192_wl3_ll", "synth", 256, 192, 3, 0, 1.0),
    ("s512_wl4_ll", "synth", 512, 512, 4, 0, 1.0),
    ("s512_wl5_ll", "synth", 512, 512, 5, 0, 1.0),      # BASELINE config 1 (outside the reference's own domain)
    ("pad200x150_wl2_ll", "synth", 200, 150, 2, 0, 1.0),  # mirror padding, ragged size
    ("zero128_wl2_ll", "zeros", 128, 128, 2, 0, 1.0),    # all-zero blocks: MSB = 32, size 1
    ("noise128_wl2_ll", "noise", 128, 128, 2, 0, 1.0),
    ("noise256_wl1_ll", "noise", 256, 256, 1, 0, 1.0),
    ("s64_wl1_ll", "synth", 64, 64, 1, 0, 1.0),          # single codeblock
    ("s320x64_wl6_ll", "synth", 320, 64, 6, 0, 1.0),     # 1-row LL, LUT holes of wl=6
    ("s256_wl3_q1", "synth", 256, 256, 3, 1, 1.0),
    ("s256_wl3_q05", "synth", 256, 256, 3, 1, 0.5),
    ("s512_wl4_q01", "synth", 512, 512, 4, 1, 0.1),
    ("noise128_wl2_q1", "noise", 128, 128, 2, 1, 1.0),
    ("pad200x150_wl3_q1", "synth", 200, 150, 3, 1, 1.0),
    ("s1024x512_wl5_q1", "synth", 1024, 512, 5, 1, 1.0),
    ("s320x128_wl6_q1", "synth", 320, 128, 6, 1, 1.0),   # level width = 2 mod 4, wl=6 LUT holes
    ("noise704x576_wl4_q03", "noise", 704, 576, 4, 1, 0.3),
]


def make_input(kind, W, H, f=0):
    return {"synth": lambda: synth.synth_frame(W, H, f), "zeros": lambda: synth.zeros_frame(W, H),
            "noise": lambda: synth.noise_frame(W, H, synth.SEED + f)}[kind]()


def u16(t):
    return t.cpu().numpy().view(np.uint16)


@pytest.fixture(scope="module")
def luts(lut_root):
    cache = {}

    def get(folder, wl):
        k = (folder, wl)
        if k not in cache:
            cache[k] = (O.lut_load(f"{lut_root}/{folder}/", wl), codec.load_lut(f"{lut_root}/{folder}/", wl))
        return cache[k]
    return get


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_stages_against_oracle(case, luts):
    name, kind, W, H, wl, lossy, qs = case
    olut, plut = luts(LUTS[(lossy, 0)], wl)
    img = make_input(kind, W, H)
    cd = codec.Codec(codec.Params(W, H, wl, bool(lossy), qs), plut)
    dev = torch.device("cuda:0")
    px = torch.from_numpy(img).to(dev)[None]

    # ---- forward DWT (+ level shift + mirror padding) ----
    opad = O.pad_mirror(img)
    if lossy:
        oc = O.dwt97_forward(opad, wl, qs)
        gc = cd.dwt_forward(px)[0].cpu().numpy()
        assert gc.dtype == np.float32
        assert np.array_equal(gc.view(np.uint32), oc.view(np.uint32)), "9/7 coefficients differ from the oracle bitwise"
        oq = O.truncate(oc)
        gq = cd.dwt_forward(px, trunc_to_int=True)
        assert np.array_equal(gq[0].cpu().numpy(), oq)
        coeff_dev = cd.dwt_forward(px)  # float path into the coder, as the reference does
    else:
        oq = O.dwt53_forward(opad, wl)
        coeff_dev = cd.dwt_forward(px)
        assert np.array_equal(coeff_dev[0].cpu().numpy(), oq), "5/3 coefficients differ"

    # ---- bit-plane coder ----
    ocs, osz = O.bpc_encode(oq, wl, olut)
    cw, sizes = cd.bpc_encode(coeff_dev)
    assert np.array_equal(sizes[0].cpu().numpy(), osz), "codeblock sizes differ"
    ref_layout = cd.export_ref_codestream(cw, sizes)[0].cpu().numpy()
    assert np.array_equal(ref_layout, ocs), "sparse codestream differs from the oracle"
    if lossy:  # integer-index input must give the same codewords as float input
        cw2, sizes2 = cd.bpc_encode(cd.dwt_forward(px, trunc_to_int=True))
        assert torch.equal(sizes2, sizes)
        assert np.array_equal(cd.export_ref_codestream(cw2, sizes2)[0].cpu().numpy(), ocs)

    # ---- bit-stream builder ----
    ohdr = O.header_pack(W, H, wl, lossy, qs, 0)
    ostream = O.bsb_pack(ocs, osz, ohdr)
    stream, off = cd.bsb_pack(cw, sizes, 0)
    off_h = off.cpu().numpy()
    assert off_h[0] == 0 and off_h[1] == ostream.size
    gstream = u16(stream[: off_h[1]])
    assert np.array_equal(gstream, ostream), "packed stream differs"

    # ---- inverse: unpack -> decode -> inverse DWT ----
    cw_b, sizes_b = cd.bsb_unpack(stream, off)
    assert torch.equal(sizes_b, sizes)
    assert np.array_equal(cd.export_ref_codestream(cw_b, sizes_b)[0].cpu().numpy(), ocs)
    q = cd.bpc_decode(cw_b, sizes_b)
    assert np.array_equal(q[0].cpu().numpy(), oq), "decoded coefficients differ"
    out = cd.dwt_inverse(q)[0].cpu().numpy()
    Wa, Ha = O.adapted_dims(W, H)
    if lossy:
        oimg = O.dwt97_inverse(oq, wl, float(np.float32(int(np.float32(qs) * 10000) / 10000.0)))
        oimg_true = O.dwt97_inverse(oq, wl, qs)
        assert np.array_equal(out, oimg_true[:H, :W]) or np.array_equal(out, oimg[:H, :W])
    else:
        assert np.array_equal(out, img), "lossless round trip is not exact"
        assert np.array_equal(O.dwt53_inverse(oq, wl)[:H, :W], img)
    cd.close()


def test_fused_frames_against_reference_golden(golden, luts):
    """pcs_encode_frames / pcs_decode_frames (host buffers) against what the reference binary wrote."""
    for name in golden.names_with_streams():
        e = golden.manifest[name]
        if e["frames"] or name == "s512_wl5_ll":  # video below; 512/wl5 is outside the reference's valid domain
            continue
        W, H, wl, lossy, qs = e["W"], e["H"], e["wl"], e["type"], e["qs"]
        _, plut = luts(e["lut"], wl)
        flags = codec.PCS_F_REF_FLAT_CROP
        cd = codec.Codec(codec.Params(W, H, wl, bool(lossy), qs, flags=flags), plut)
        img = golden.input(name)
        stream, lens = cd.encode_frames(img[None])
        ref = golden.stream(name)
        assert lens[0] == ref.size, name
        assert np.array_equal(stream, ref), f"{name}: codestream differs from the reference binary's"
        dec = cd.decode_frames(ref, np.array([ref.size], dtype=np.int32))[0].reshape(-1)
        refdec = golden.decoded(name)
        if refdec is None:  # lossless, reference round-tripped: decoded == input
            assert e["dec_equal_input"]
            assert np.array_equal(dec, img.reshape(-1)), name
        elif not lossy:     # padded case: compare with the reference's (flat-cropped) output
            assert np.array_equal(dec, refdec), name
        else:
            d = np.abs(dec.astype(np.int32) - refdec.astype(np.int32))
            assert d.max() <= 1 and (d != 0).mean() <= 1e-4, (name, int(d.max()), float((d != 0).mean()))
            mse_a = np.mean((dec.astype(np.float64) - img.reshape(-1)) ** 2)
            mse_b = np.mean((refdec.astype(np.float64) - img.reshape(-1)) ** 2)
            assert abs(10 * np.log10(mse_b / mse_a)) <= 0.01, name  # PSNR within 0.01 dB
        cd.close()


def test_video_batches_against_reference_golden(golden, luts):
    for name in ("v128_f4_wl2_ll", "v256_f3_wl3_q1"):
        e = golden.manifest[name]
        W, H, F, wl, lossy, qs = e["W"], e["H"], e["frames"], e["wl"], e["type"], e["qs"]
        _, plut = luts(e["lut"], wl)
        ref = golden.stream(name)
        ref_lens = np.array([int(t) for t in e["size_text"].split(",")], dtype=np.int32)
        frames = synth.synth_video(W, H, F)
        for max_batch in (1, 3, F):
            cd = codec.Codec(codec.Params(W, H, wl, bool(lossy), qs, frames=F, max_batch=max_batch), plut)
            stream, lens = cd.encode_frames(frames)
            assert np.array_equal(lens, ref_lens), (name, max_batch)
            assert np.array_equal(stream, ref), (name, max_batch)
            dec = cd.decode_frames(ref, ref_lens)
            if lossy:
                refdec = golden.decoded(name).reshape(F, H, W)
                d = np.abs(dec.astype(np.int32) - refdec.astype(np.int32))
                assert d.max() <= 1 and (d != 0).mean() <= 1e-4
            else:
                assert np.array_equal(dec, frames)
            cd.close()


def test_reference_coefficients(luts):
    """Forward 9/7 coefficients against the reference's own DWTEncode output (oracle/ref_harness dump)."""
    import os
    p = os.path.join(os.path.dirname(__file__), "golden", "ref_coef.npz")
    z = np.load(p)
    _, plut = luts("video_lossy", 3)
    cd = codec.Codec(codec.Params(256, 256, 3, True, 1.0, max_batch=3), plut)
    px = torch.from_numpy(synth.synth_video(256, 256, 3)).cuda()
    c = cd.dwt_forward(px).cpu().numpy()
    for f in range(3):
        ref = z[f"v256_f3_wl3_q1.f{f}"]
        rel = np.abs(c[f] - ref) / np.maximum(np.abs(ref), 1.0)  # relative error (absolute below |c| = 1)
        # north_star tolerance 1e-4.  The reference binary is itself not deterministic at this level: rows at
        # the seam of its bottom warp tile are written by two warps with different FMA contractions (a write
        # race, SURVEY note / DESIGN.md section 2) and the 1-ulp differences of level-1/2 LL values (~2e-4 at
        # magnitude 2048) propagate into the next level.  Measured: <= 4 samples of 65536 per frame between
        # 1.0e-4 and 1.6e-4, all on those rows; everything else is inside 1e-4 and >= 99 % is bit-identical.
        assert (rel > 1e-4).mean() <= 1e-4 and rel.max() <= 2.5e-4, (float(rel.max()), int((rel > 1e-4).sum()))
        assert (c[f].view(np.uint32) != ref.view(np.uint32)).mean() <= 2e-2
        # quantisation indices (int)c: differ by at most 1 on at most 1e-5 of the samples
        qa, qb = c[f].astype(np.int32), ref.astype(np.int32)
        assert np.abs(qa - qb).max() <= 1 and (qa != qb).mean() <= 1e-5 + 1.0 / qa.size
    cd.close()


def test_full_size_properties(luts):
    """BASELINE.json-sized inputs: size-independent properties + sha256 of the reference's stream."""
    import hashlib
    import json
    import os
    man = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_manifest.json")))
    # config 2 lossless twin and config 3: byte-identical to the reference binary (sha256 of its file)
    for name in ("cfg2_4096_wl5_ll", "cfg3_8k_wl6_ll"):
        e = man[name]
        _, plut = luts(e["lut"], e["wl"])
        cd = codec.Codec(codec.Params(e["W"], e["H"], e["wl"], False, 1.0), plut)
        img = synth.synth_frame(e["W"], e["H"])
        stream, lens = cd.encode_frames(img[None])
        assert stream.size * 2 == e["enc_len"]
        assert hashlib.sha256(stream.tobytes()).hexdigest() == e["enc_sha256"], name
        dec = cd.decode_frames(stream, lens)[0]
        assert np.array_equal(dec, img)  # encode -> decode round trip is the identity
        cd.close()
    # config 2 (lossy, the round-1 headline configuration): byte-identical to the reference binary as well.  The reference is
    # deterministic on this input (five runs per qs with one sha256 each: tests/golden/ref_runs/cfg2_determinism.json),
    # and the forward 9/7 is matched bit for bit, so the sha256 of its stream is asserted; decoded pixels: PSNR equal
    # to the reference's within 0.01 dB (its inverse uses -use_fast_math divisions).
    for name in ("cfg2_4096_wl5_q1", "cfg2_4096_wl5_q05", "cfg2_4096_wl5_q01"):
        e = man[name]
        _, plut = luts(e["lut"], e["wl"])
        cd = codec.Codec(codec.Params(e["W"], e["H"], e["wl"], True, e["qs"]), plut)
        img = synth.synth_frame(e["W"], e["H"])
        stream, lens = cd.encode_frames(img[None])
        assert stream.size * 2 == e["enc_len"], (name, stream.size * 2, e["enc_len"])
        assert hashlib.sha256(stream.tobytes()).hexdigest() == e["enc_sha256"], name
        dec = cd.decode_frames(stream, lens)[0]
        mse = np.mean((dec.astype(np.float64) - img) ** 2)
        psnr = 10 * np.log10(255.0 ** 2 / mse)
        assert abs(psnr - e["dec_psnr"]) <= 0.01, (name, psnr, e["dec_psnr"])
        cd.close()


def test_config4_video_against_reference(luts):
    """BASELINE config 4 shape (3840x2160 lossy video, mirrored rows 2160..2175): the first 4 frames byte for byte
    against the reference binary (sha256 of its stream and its `_SIZE` text), then a 16-frame batch whose frames
    must be independent of how they are batched (frame f of the batch == frame f coded alone, header aside)."""
    import hashlib
    import json
    import os
    man = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_manifest.json")))
    e = man["cfg4_4k_f4_wl5_q1"]
    W, H, wl = e["W"], e["H"], e["wl"]
    _, plut = luts(e["lut"], wl)
    frames = synth.synth_video(W, H, 16)
    cd = codec.Codec(codec.Params(W, H, wl, True, e["qs"], frames=4, max_batch=4), plut)
    stream, lens = cd.encode_frames(frames[:4])
    assert ",".join(str(int(v)) for v in lens) == e["size_text"]
    assert hashlib.sha256(stream.tobytes()).hexdigest() == e["enc_sha256"]
    dec = cd.decode_frames(stream, lens)
    assert hashlib.sha256(dec.tobytes()).hexdigest() == e["dec_sha256"] or \
        (np.abs(dec.astype(np.int32) - frames[:4]).max() <= e["dec_maxabs"] + 1)
    mse = np.mean((dec.astype(np.float64) - frames[:4]) ** 2)
    assert abs(10 * np.log10(255.0 ** 2 / mse) - e["dec_psnr"]) <= 0.01
    cd.close()
    # batching invariance over 16 frames with a different batch size (frames > 0 carry 0xFFFF x 9 instead of the header)
    cd = codec.Codec(codec.Params(W, H, wl, True, e["qs"], frames=4, max_batch=16), plut)
    s16, l16 = cd.encode_frames(frames)
    assert np.array_equal(l16[:4], lens) and np.array_equal(s16[: stream.size], stream)
    off = np.concatenate([[0], np.cumsum(l16)])
    one = codec.Codec(codec.Params(W, H, wl, True, e["qs"], frames=4, max_batch=1), plut)
    for f in (5, 15):
        sf, lf = one.encode_frames(frames[f:f + 1], first_frame_index=f)
        assert lf[0] == l16[f] and np.array_equal(sf, s16[off[f]:off[f + 1]])
    d16 = cd.decode_frames(s16, l16)
    assert np.array_equal(d16[:4], dec)
    assert np.abs(d16.astype(np.int32) - frames).max() <= 3
    one.close(); cd.close()


def test_config5_row_bands(luts):
    """BASELINE config 5 shape: a 32768-wide still cut into 32768x4096 row bands (one band == one reference
    'frame'); two bands, lossless.  The reference binary itself fails on this size (tests/golden manifest:
    rc_dec = -8), so the checks are the size-independent ones: exact round trip, band independence, and the
    packed layout invariants (header only on global band 0, per-block sizes sum to the band length)."""
    W, H, wl = 32768, 4096, 5
    _, plut = luts("n1_lossless", wl)
    bands = np.stack([synth.synth_frame(W, H, 0, 0, H * b) for b in range(2)])
    cd = codec.Codec(codec.Params(W, H, wl, False, 1.0, frames=8, max_batch=2), plut)
    stream, lens = cd.encode_frames(bands)
    ncb = cd.ncb
    off = np.concatenate([[0], np.cumsum(lens)])
    for b in range(2):
        fr = stream[off[b]:off[b + 1]]
        sizes = fr[10:10 + 2 * ncb:2].astype(np.int64)
        assert sizes.sum() + ncb + 10 == lens[b]          # HTotalBSSize = sum(size) + nCB + 10
        assert fr[-1] == 0xFFFF
        assert (fr[:9] == 0xFFFF).all() == (b != 0)       # only global band 0 carries the header
    dec = cd.decode_frames(stream, lens)
    assert np.array_equal(dec, bands)
    # band 1 coded alone (as rank 1 of a sharded job would) gives the same bytes
    alone, l1 = cd.encode_frames(bands[1:2], first_frame_index=1)
    assert l1[0] == lens[1] and np.array_equal(alone, stream[off[1]:off[2]])
    cd.close()


@pytest.mark.parametrize("lossy", [0, 1])
def test_batched_dwt_matches_single_frames(lossy, luts):
    """The batched DWT drivers overlap level 1 with the deeper levels of the previous frame group on auxiliary
    streams; a batch must give, frame by frame, exactly what single-frame calls give (forward and inverse)."""
    W, H, wl, B = 512, 384, 4, 7
    _, plut = luts(LUTS[(lossy, 1)], wl)
    cd = codec.Codec(codec.Params(W, H, wl, bool(lossy), 0.7, max_batch=B), plut)
    px = torch.from_numpy(synth.synth_video(W, H, B)).cuda()
    for it in range(3):  # repeated calls reuse the fork/join events
        c_all = cd.dwt_forward(px, trunc_to_int=True)
        c_one = torch.cat([cd.dwt_forward(px[f:f + 1], trunc_to_int=True) for f in range(B)])
        assert torch.equal(c_all, c_one)
        p_all = cd.dwt_inverse(c_all)
        p_one = torch.cat([cd.dwt_inverse(c_all[f:f + 1]) for f in range(B)])
        assert torch.equal(p_all, p_one)
    if not lossy:
        assert torch.equal(p_all, px)
    cd.close()


def test_decoder_survives_malformed_streams(luts):
    """Hostile / damaged codestreams must not crash, hang or fault the GPU: the decoder clamps every size and slot it
    reads from the stream (bsb_unpack_kernel, dec_step) and frame lengths are validated on the host.  The decoded
    pixels are garbage by definition; what is checked is that every call returns and the context stays usable."""
    W, H, wl = 256, 192, 3
    _, plut = luts("n1_lossless", wl)
    cd = codec.Codec(codec.Params(W, H, wl, False, 1.0, max_batch=2), plut)
    img = synth.synth_frame(W, H)
    good, lens = cd.encode_frames(img[None])
    rng = np.random.default_rng(7)
    for trial in range(12):
        bad = good.copy()
        kind = trial % 4
        if kind == 0:    # random words everywhere behind the header
            bad[9:] = rng.integers(0, 65536, bad.size - 9, dtype=np.uint16)
        elif kind == 1:  # corrupt the per-block (MSB, size) table only
            n = 2 * cd.ncb
            bad[9:9 + n] = rng.integers(0, 65536, n, dtype=np.uint16)
        elif kind == 2:  # flip a few bits in the codewords
            idx = rng.integers(9 + 2 * cd.ncb, bad.size, 64)
            bad[idx] ^= np.uint16(1) << rng.integers(0, 16, 64).astype(np.uint16)
        else:            # all-ones sizes
            bad[10:10 + 2 * cd.ncb:2] = 0xFFFF
        out = cd.decode_frames(bad, lens)
        assert out.shape == (1, H, W)
    # lengths that cannot be right are rejected on the host with an error code, not passed to the GPU
    with pytest.raises(codec.PcsError):
        cd.decode_frames(good, np.array([5], dtype=np.int32))
    with pytest.raises(codec.PcsError):
        cd.decode_frames(good, np.array([2 ** 30], dtype=np.int32))
    # the context still works
    assert np.array_equal(cd.decode_frames(good, lens)[0], img)
    cd.close()


_FUSE_SNIPPET = r"""
import hashlib, sys
sys.path[:0] = [{root!r}, {pkg!r}]
import numpy as np
import torch
from picsong_b200 import codec, synth
out = []
# W, H, wl, lossy, qs, frames: several CTAs across (W > 928), partial last CTA, padded heights, wl == 2 (LL2 is final), 5/3 and 9/7
for (W, H, wl, lossy, qs, B) in [(512, 384, 3, 0, 1.0, 3), (1024, 512, 4, 1, 1.0, 2), (256, 192, 2, 1, 0.5, 1), (768, 320, 5, 0, 1.0, 2),
                                 (1984, 200, 2, 0, 1.0, 1), (3840, 1100, 5, 1, 1.0, 1), (2048, 2048, 5, 1, 0.3, 1), (960, 1000, 3, 0, 1.0, 2),
                                 (4096, 640, 5, 1, 1.0, 2), (64, 64, 2, 1, 1.0, 1), (128, 4096, 3, 0, 1.0, 1)]:
    folder = "video_lossy" if lossy else "video_lossless"
    cd = codec.Codec(codec.Params(W, H, wl, bool(lossy), qs, max_batch=B), codec.load_lut({lut_root!r} + "/" + folder + "/", wl))
    px = torch.from_numpy(np.stack([synth.synth_frame(W, H, f) for f in range(B)])).cuda()
    for trunc in (True, False):
        c = cd.dwt_forward(px, trunc_to_int=trunc)
        out.append(hashlib.sha256(c.cpu().numpy().tobytes()).hexdigest())
    s, lens = cd.encode_frames(px.cpu().numpy(), 0)
    out.append(hashlib.sha256(np.asarray(s).tobytes()).hexdigest())
    back = cd.decode_frames(s, lens)
    if not lossy:
        assert np.array_equal(back, px.cpu().numpy())
    out.append(hashlib.sha256(np.asarray(back).tobytes()).hexdigest())
print("HASHES " + " ".join(out))
"""


def test_fused_levels_match_per_level_launches(lut_root):
    """Levels 1+2 of the forward transform in one launch (dwt_fused.cu: LL1 handed from producer to consumer warps through
    shared memory, TMA row loads; the default) against one launch per level (PCS_DWT_FUSE=1): same coefficients, same
    streams and same decoded pixels bit for bit, also with short chunks (PCS_FUSED_RP) so that chunk seams and the LL1 ring
    wrap many times.  The switches are read once per
    process, hence the subprocesses."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = _FUSE_SNIPPET.format(root=root, pkg=os.path.join(root, "cuda-image-and-video-codec_b200"), lut_root=str(lut_root))
    res = {}
    for name, extra in (("per_level", {"PCS_DWT_FUSE": "1"}), ("fused", {}), ("fused_rp4", {"PCS_FUSED_RP": "4"}), ("fused_rp20", {"PCS_FUSED_RP": "20"})):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **extra), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res[name] = [ln for ln in r.stdout.splitlines() if ln.startswith("HASHES ")][0].split()
    for name in ("fused", "fused_rp4", "fused_rp20"):
        bad = [i for i, (x, y) in enumerate(zip(res["per_level"], res[name])) if x != y]
        assert not bad, (name, bad)


_PIPE_SNIPPET = r"""
import hashlib, sys
sys.path[:0] = [{root!r}, {pkg!r}]
import numpy as np
from picsong_b200 import codec, synth
out = []
for (W, H, wl, lossy, B, n) in [(256, 192, 3, 0, 16, 16), (320, 256, 2, 1, 16, 13), (256, 256, 2, 0, 8, 5)]:
    folder = "video_lossy" if lossy else "video_lossless"
    cd = codec.Codec(codec.Params(W, H, wl, bool(lossy), 1.0, frames=n, max_batch=B), codec.load_lut({lut_root!r} + "/" + folder + "/", wl))
    px = np.stack([synth.synth_frame(W, H, f) for f in range(n)])
    s, lens = cd.encode_frames(px, 0)
    back = cd.decode_frames(s, lens)
    if not lossy:
        assert np.array_equal(back, px)
    out += [hashlib.sha256(np.asarray(s).tobytes()).hexdigest(), hashlib.sha256(lens.tobytes()).hexdigest(), hashlib.sha256(back.tobytes()).hexdigest()]
print("HASHES " + " ".join(out))
"""


def test_host_buffer_pipeline_chunking_is_transparent(lut_root):
    """pcs_encode_frames / pcs_decode_frames cut a call into chunks on four slot streams (tapered first / last chunks,
    at most two chunks' kernels in flight).  Small test frames normally travel as ONE chunk, so the chunked schedule
    is forced here (PCS_PIPE_CHUNK, read once per process -> subprocesses) and must not change a single byte."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = _PIPE_SNIPPET.format(root=root, pkg=os.path.join(root, "cuda-image-and-video-codec_b200"), lut_root=str(lut_root))
    res = {}
    for name, extra in (("one_chunk", {}), ("chunk2", {"PCS_PIPE_CHUNK": "2"}), ("chunk1_depth3", {"PCS_PIPE_CHUNK": "1", "PCS_PIPE_DEPTH": "3"}),
                        ("chunk2_notaper_lag1", {"PCS_PIPE_CHUNK": "2", "PCS_PIPE_TAPER": "0", "PCS_PIPE_LAG": "1"})):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **extra), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        res[name] = [ln for ln in r.stdout.splitlines() if ln.startswith("HASHES ")][0]
    assert len(set(res.values())) == 1, res
