"""CPU tests: the oracle (oracle/picsong_oracle.c) against the outputs of the UNMODIFIED reference binary
run on a B200 (tests/golden, produced by oracle/run_ref_golden.py + oracle/make_golden.py), plus the
known-answer vectors and identities SURVEY.md derives from the reference source."""
import hashlib

import numpy as np
import pytest

from oracle import oracle as O
from picsong_b200 import synth


def _lut(lut_root, folder, wl):
    return O.lut_load(f"{lut_root}/{folder}/", wl)


def test_oracle_reproduces_reference_streams(golden, lut_root):
    """Every still the reference coded on the B200: the oracle's codestream is byte-identical."""
    checked = 0
    for name in golden.names_with_streams():
        e = golden.manifest[name]
        if e["frames"] or name == "s512_wl5_ll":
            continue
        lut = _lut(lut_root, e["lut"], e["wl"])
        s = O.encode_frame(golden.input(name), e["wl"], bool(e["type"]), e["qs"], lut)
        ref = golden.stream(name)
        assert s.size == ref.size and np.array_equal(s, ref), name
        assert hashlib.sha256(s.tobytes()).hexdigest() == e["enc_sha256"]
        checked += 1
    assert checked >= 10


def test_oracle_reproduces_reference_video(golden, lut_root):
    for name in ("v128_f4_wl2_ll", "v256_f3_wl3_q1"):
        e = golden.manifest[name]
        lut = _lut(lut_root, e["lut"], e["wl"])
        ref, off = golden.stream(name), 0
        lens = [int(t) for t in e["size_text"].split(",")]
        for f in range(e["frames"]):
            s = O.encode_frame(synth.synth_frame(e["W"], e["H"], f), e["wl"], bool(e["type"]), e["qs"], lut,
                               first=(f == 0), frames=e["frames"])
            assert np.array_equal(s, ref[off:off + lens[f]]), (name, f)
            if f > 0:
                assert np.all(s[:9] == 0xFFFF)  # only frame 0 carries the header (BitStreamBuilder.cu:276-278)
            off += lens[f]
        assert off == ref.size


def test_oracle_decodes_reference_streams(golden, lut_root):
    for name in golden.names_with_streams():
        e = golden.manifest[name]
        if e["frames"] or name == "s512_wl5_ll":
            continue
        W, H, wl, lossy = e["W"], e["H"], e["wl"], bool(e["type"])
        lut = _lut(lut_root, e["lut"], wl)
        ref = golden.stream(name)
        b = O.header_unpack(ref[:9])
        assert (b[0], b[10], b[4], b[6]) == (W * H, H, wl, int(lossy))
        Wa, Ha = O.adapted_dims(W, H)
        px = O.decode_frame(ref, Wa, Ha, wl, lossy, b[7] / 10000.0, lut)
        flat = px.reshape(-1)[: W * H]  # the reference's output quirk (IOManager.ipp:332-335)
        refdec = golden.decoded(name)
        if refdec is None:
            assert np.array_equal(px[:H, :W], golden.input(name)), name
        elif not lossy:
            assert np.array_equal(flat, refdec), name
        else:  # inverse 9/7 uses fast-math divisions in the reference: +-1 on a handful of pixels
            d = np.abs(flat.astype(int) - refdec.astype(int))
            assert d.max() <= 1 and (d != 0).mean() <= 1e-4, name


def test_config1_is_outside_the_reference_domain(golden):
    """512x512 wl=5: the reference's own decode of its own stream is NOT the input (SURVEY note 4)."""
    e = golden.manifest["s512_wl5_ll"]
    assert e["rc_enc"] == 0 and e["rc_dec"] == 0 and e["dec_equal_input"] is False and e["dec_psnr"] < 30


def test_full_size_hashes(lut_root, golden):
    """BASELINE config 2 (lossless twin): oracle == reference on the full 4096x4096 frame (sha256)."""
    e = golden.manifest["cfg2_4096_wl5_ll"]
    lut = _lut(lut_root, e["lut"], e["wl"])
    # bounded: code the first 1024 rows' worth of codeblocks only would change the stream; do the full frame (~6 s)
    s = O.encode_frame(synth.synth_frame(4096, 4096), 5, False, 1.0, lut)
    assert s.size * 2 == e["enc_len"]
    assert hashlib.sha256(s.tobytes()).hexdigest() == e["enc_sha256"]


def test_lut_loader(lut_root):
    """Annex D: 2730 / 3360 / 3990 ints for wl = 4 / 5 / 6 with 0 / 0 / 630 zero entries, max 127."""
    for wl, n, zeros in ((4, 2730, 0), (5, 3360, 0), (6, 3990, 630)):
        lut = _lut(lut_root, "n1_lossless", wl)
        assert lut.table.size == n and int((lut.table == 0).sum()) == zeros
        assert lut.table.max() == 127 and lut.table.min() >= 0
    assert list(_lut(lut_root, "n1_lossy", 5).hdr) == [15, 3, 1, 4, 9, 7, 3, 15]


def test_header_known_answers():
    """Annex C.1 vectors derived from BitStreamBuilder.cpp:35-94."""
    kat = [
        ((512, 512, 5, 0, 1.0, 0), "0000 0004 c024 8042 0338 0000 1002 0000"),
        ((4096, 4096, 5, 1, 0.5, 0), "0000 0100 c024 4442 029c 0000 1010 0000"),
        ((7680, 4320, 6, 0, 1.0, 0), "4000 01fa 4024 8043 0338 e000 1010 0000"),
        ((3840, 2160, 5, 1, 1.0, 600), "9000 007e c024 8442 0338 7000 1008 012c"),
        ((32768, 4096, 5, 0, 1.0, 8), "0000 0800 c024 8042 0338 0000 1010 0004"),
    ]
    for (W, H, wl, lossy, qs, frames), want in kat:
        h = O.header_pack(W, H, wl, lossy, qs, frames)
        assert " ".join(f"{v:04x}" for v in h[:8]) == want and h[8] == 0
        b = O.header_unpack(h)
        assert (b[0] // b[10], b[10], b[4], b[6], b[14]) == (W, H, wl, lossy, frames)
        assert b[7] == int(np.float32(qs) * 10000)


def test_bsb_literal_equals_closed_form():
    """The reference's per-codeword binary search (restated literally) == the closed-form layout."""
    rng = np.random.default_rng(7)
    for n in (1, 2, 7, 64, 300, 2040):
        for kind in range(3):
            sizes = {0: rng.integers(1, 4097, n), 1: np.ones(n, dtype=np.int64), 2: rng.integers(3800, 4097, n)}[kind].astype(np.int32)
            cs = rng.integers(0, 65536, n * 4096).astype(np.int32)
            hdr = rng.integers(0, 65536, 9).astype(np.uint16)
            a = O.bsb_pack(cs, sizes, hdr)
            b = O.bsb_pack(cs, sizes, hdr, closed=True)
            assert np.array_equal(a, b)
            assert a.size == sizes.sum() + n + 10 and a[-1] == 0xFFFF
            cs2, sz2, tot = O.bsb_unpack(a, n)
            assert np.array_equal(sz2, sizes) and tot == a.size
            for cb in range(n):
                assert np.array_equal(cs2[cb * 4096: cb * 4096 + sizes[cb]], cs[cb * 4096: cb * 4096 + sizes[cb]] & 0xFFFF)
                assert np.all(cs2[cb * 4096 + sizes[cb]: (cb + 1) * 4096] == -1)


def test_lifting_identities():
    rng = np.random.default_rng(3)
    for (H, W, wl) in ((64, 64, 1), (128, 192, 3), (64, 320, 6)):
        px = rng.integers(0, 256, (H, W)).astype(np.uint8)
        assert np.array_equal(O.dwt53_inverse(O.dwt53_forward(px, wl), wl), px)  # perfect reconstruction
    px = synth.synth_frame(256, 256)
    c = O.dwt97_forward(px, 3, 1.0)
    rec = O.dwt97_inverse(O.truncate(c), 3, 1.0)
    mse = np.mean((rec.astype(float) - px) ** 2)
    assert 10 * np.log10(255 ** 2 / mse) > 50
    # constant image: all energy in LL, high bands exactly 0
    flat = np.full((128, 128), 200, dtype=np.uint8)
    c = O.dwt53_forward(flat, 2)
    assert np.all(c[:32, :32] == 72) and np.count_nonzero(c) == 32 * 32


def test_padding_is_edge_inclusive_mirror():
    img = np.arange(150 * 200, dtype=np.uint32).reshape(150, 200).astype(np.uint8)
    p = O.pad_mirror(img)
    assert p.shape == (192, 256)
    assert np.array_equal(p[:150, 200:], img[:, 199:143:-1])
    assert np.array_equal(p[150:], p[149:107:-1])


def test_bpc_round_trip_and_edge_cases(lut_root):
    lut = _lut(lut_root, "n1_lossless", 2)
    rng = np.random.default_rng(11)
    zero = np.zeros((128, 128), dtype=np.int32)
    cs, sz = O.bpc_encode(zero, 2, lut)
    assert np.all(sz == 1) and np.all(cs[::4096] == 32)  # MSB = 32, nothing coded
    assert np.array_equal(O.bpc_decode(cs, sz, 128, 128, 2, lut), zero)
    for scale in (3, 200, 3000):  # |q| < 2^15: also exact through the raw-dump fallback (15 magnitude bits kept)
        q = np.clip(rng.laplace(0, scale, (128, 128)), -32767, 32767).astype(np.int32)
        cs, sz = O.bpc_encode(q, 2, lut)
        assert np.array_equal(O.bpc_decode(cs, sz, 128, 128, 2, lut), q), scale
    # expansion fallback (BPCEngine.cu:2015-2024): >4095 codewords -> raw 16-bit dump, size 4096
    q = rng.integers(-16000, 16000, (64, 64)).astype(np.int32)
    lut1 = _lut(lut_root, "n1_lossless", 1)
    cs, sz = O.bpc_encode(q, 1, lut1)
    assert sz[0] == 4096
    assert np.array_equal(O.bpc_decode(cs, sz, 64, 64, 1, lut1), q)


def test_lane_classification_pins_block0(golden, lut_root):
    """findSubband is evaluated per lane (BPCEngine.cu:1975,1990): block 0 of an image whose LL band is
    narrower than 64 mixes LL and HL tables.  Pinned by the reference stream of s128_wl2_ll."""
    e = golden.manifest["s128_wl2_ll"]
    ref = golden.stream("s128_wl2_ll")
    lut = _lut(lut_root, e["lut"], 2)
    q = O.dwt53_forward(O.pad_mirror(golden.input("s128_wl2_ll")), 2)
    cs, sz = O.bpc_encode(q, 2, lut)
    assert sz[0] == ref[10]


RGB_CASES = ["rgb128_wl2_ll", "rgb256x192_wl3_ll", "rgb256x192_wl3_q1", "rgb200x150_wl2_q05", "rgbv128_f3_wl2_ll", "rgbv256_f2_wl3_q1"]


@pytest.mark.parametrize("name", RGB_CASES)
def test_oracle_rgb_matches_reference_binary(name, golden, lut_root):
    """RGB path of the oracle (colour transform, three component streams with the R/G/B tables, header placement)
    against the unmodified reference binary run with -isRGB 1: streams and _SIZE values byte for byte."""
    from picsong_b200 import synth
    e = golden.manifest[name]
    W, H, wl, lossy, qs, F = e["W"], e["H"], e["wl"], bool(e["type"]), e["qs"], e["frames"]
    luts = [O.lut_load(f"{lut_root}/{e['lut']}/", wl, c) for c in range(3)]
    sizes = [int(t) for t in e["size_text"].split(",")]
    out = []
    for f in range(max(F, 1)):
        rgb = synth.synth_rgb(W, H, f if F else 0)
        out += O.encode_rgb_frame(rgb, wl, lossy, qs, luts, [True, False, False] if not F else [f == 0] * 3, frames=F)
    assert [s.size for s in out] == sizes
    assert np.array_equal(np.concatenate(out), golden.stream(name))
    Wa, Ha = O.adapted_dims(W, H)
    off = np.concatenate([[0], np.cumsum(sizes)])
    ref = golden.stream(name)
    for f in range(max(F, 1)):
        d = O.decode_rgb_frame([ref[off[3 * f + c]:off[3 * f + c + 1]] for c in range(3)], Wa, Ha, wl, lossy, qs, luts)[:, :H, :W]
        src = synth.synth_rgb(W, H, f if F else 0)
        if not lossy:
            assert np.array_equal(d, src)
        elif W % 64 == 0:
            r = golden.decoded(name).reshape(max(F, 1), 3, H, W)[f]
            diff = np.abs(d.astype(np.int32) - r.astype(np.int32))
            assert diff.max() <= 1 and (diff != 0).mean() <= 1e-4


def test_oracle_matches_reference_coder_on_rare_branches(lut_root):
    """Expansion fallback (size 4096) and bit planes >= 16: the oracle's coder + bit-stream builder and its decoder against
    the reference's own BPCCuda<int>::Code / ::Decode run on the synthetic planes of tests/coef_planes.py
    (tests/golden/coef_planes.npz, made with `ref_harness ... coef=`, see tests/golden/ref_runs/README.md)."""
    import os

    import coef_planes
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "coef_planes.npz"))
    lut = O.lut_load(f"{lut_root}/n1_lossless/", coef_planes.WL)
    W, H, wl = coef_planes.W, coef_planes.H, coef_planes.WL
    seen_expansion = seen_high = False
    for name in coef_planes.CASES:
        q = coef_planes.plane(name)
        cs, sz = O.bpc_encode(q, wl, lut)
        dq = O.bpc_decode(cs, sz, W, H, wl, lut)
        if f"{name}.stream" not in g:      # every block expands: the reference itself fails on these (ref_runs/README.md)
            assert (sz == 4096).all()
            continue
        assert np.array_equal(O.bsb_pack(cs, sz, O.header_pack(W, H, wl, 0, 1.0, 0)), g[f"{name}.stream"]), name
        assert np.array_equal(dq, g[f"{name}.q"]), name
        seen_expansion |= bool((sz == 4096).any())
        seen_high |= bool(np.abs(q.astype(np.int64)).max() >= 1 << 16 and (sz < 4096).any())
    assert seen_expansion and seen_high


F4_CASES = [("bps16_ll", 0, 16, 0), ("bps12_ll", 0, 12, 0), ("signed_ll", 0, 8, 1), ("signed_q1", 1, 8, 1), ("bps16_q1", 1, 16, 0)]


def test_oracle_bit_depth_and_signed_input_match_reference(lut_root):
    """-bps 16 / 12 and -signedOrUnsigned 1: only the level shift changes (2^(bps-1), or 0 for signed
    samples).  Oracle == the reference executable's streams and decoded pixels (tests/golden/f4_ref.npz); for the lossy
    -bps 16 run within the 9/7 tolerance (see tests/golden/ref_runs/README.md)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "f4_ref.npz"))
    img = g["input"]
    H, W = img.shape
    try:
        for name, lossy, bps, signed in F4_CASES:
            lut = O.lut_load(f"{lut_root}/{'n1_lossy' if lossy else 'n1_lossless'}/", 3)
            O.set_sample_offset(0 if signed else 1 << (bps - 1))
            px = O.pad_mirror(img)
            q = O.truncate(O.dwt97_forward(px, 3, 1.0)) if lossy else O.dwt53_forward(px, 3)
            cs, sz = O.bpc_encode(q, 3, lut)
            st = O.bsb_pack(cs, sz, O.header_pack(W, H, 3, lossy, 1.0, 0, bitDepth=bps, bps=bps, isSigned=signed))
            ref = g[name + ".stream"]
            assert np.array_equal(st[:9], ref[:9]), name                      # header words carry bps / signedness
            if name != "bps16_q1":
                assert np.array_equal(st, ref), name
                assert np.array_equal(O.decode_frame(ref, W, H, 3, bool(lossy), 1.0, lut), g[name + ".dec"]), name
            else:
                rcs, rsz, _ = O.bsb_unpack(ref, (W // 64) * (H // 64))
                rq = O.bpc_decode(rcs, rsz, W, H, 3, lut)
                assert np.abs(rq - q).max() <= 1 and (rq != q).mean() <= 1e-4
                # pixels: the reference's -use_fast_math divisions at magnitude 32768 (one fp32 ulp = 0.004) move ~0.7 % of the
                # roundings by one grey level; PSNR against the input is the north_star criterion (within 0.01 dB)
                mine, theirs = O.decode_frame(ref, W, H, 3, True, 1.0, lut), g[name + ".dec"]
                d = np.abs(mine.astype(int) - theirs.astype(int))
                assert d.max() <= 1 and (d != 0).mean() <= 1e-2
                psnr = lambda a: 10 * np.log10(255.0 ** 2 / np.mean((a.astype(np.float64) - img) ** 2))  # noqa: E731
                assert abs(psnr(mine) - psnr(theirs)) <= 0.01
            if not lossy:
                assert np.array_equal(g[name + ".dec"], img)
    finally:
        O.set_sample_offset(128)


def test_oracle_64x32_codeblocks_match_reference_build(lut_root):
    """64 x 32 codeblocks (the reference rebuilt with -DCBLOCK_LENGTH=32, BASELINE configs[2]'s sweep point): oracle ==
    that build's streams byte for byte (tests/golden/cb32_ref.npz, scripts/r2_gpu_i.sh)."""
    import os

    from picsong_b200 import synth
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cb32_ref.npz"))
    cases = [("cb32_s256x160_wl3_ll", "synth", 256, 160, 3, 0, 1.0), ("cb32_s512x288_wl3_q1", "synth", 512, 288, 3, 1, 1.0),
             ("cb32_pad200x150_wl2_ll", "synth", 200, 150, 2, 0, 1.0), ("cb32_noise128x96_wl2_ll", "noise", 128, 96, 2, 0, 1.0),
             ("cb32_s1024x544_wl4_q05", "synth", 1024, 544, 4, 1, 0.5)]
    try:
        O.set_cb_rows(32)
        for name, kind, W, H, wl, lossy, qs in cases:
            lut = O.lut_load(f"{lut_root}/{'n1_lossy' if lossy else 'n1_lossless'}/", wl)
            img = synth.synth_frame(W, H) if kind == "synth" else synth.noise_frame(W, H)
            assert O.adapted_dims(W, H) == (-(-W // 64) * 64, -(-H // 32) * 32)
            s = O.encode_frame(img, wl, bool(lossy), qs, lut)
            assert np.array_equal(s, g[name + ".stream"]), name
            Wa, Ha = O.adapted_dims(W, H)
            dec = O.decode_frame(s, Wa, Ha, wl, bool(lossy), qs, lut)[:H, :W]
            if not lossy:
                assert np.array_equal(dec, img), name
            elif name + ".dec" in g:
                d = np.abs(dec.astype(int) - g[name + ".dec"].reshape(H, W).astype(int))
                assert d.max() <= 1 and (d != 0).mean() <= 1e-4
    finally:
        O.set_cb_rows(64)
