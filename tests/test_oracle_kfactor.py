"""CPU test: the oracle's restatement of the reference's complexity-scalable ("bulk", -k > 0) ENCODER
(BPC/BPCEngine.cu:1285-1662,1683-1716) against streams written by the unmodified reference binary on a B200
(tests/golden/k_factor, made by oracle/run_ref_golden.py --k; LUT tables: tests/golden/lut_tables_k.npz, made by
oracle/make_golden_k.py).  The new CUDA code does not implement this mode yet (PCS_E_UNSUPPORTED): this pins the
oracle that has to exist first.

Byte-exact for every geometry in which no codeblock straddles a subband.  Where one does (block 0 when the deepest
subbands are narrower than 64), the lanes of that warp leave the reference's plane loop at different planes, which
is undefined behaviour there -- its own decoder aborts on those streams -- so block 0 is exempt and all other
blocks must still match."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from picsong_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
KDIR = os.path.join(HERE, "golden", "k_factor")
MAN = json.load(open(os.path.join(KDIR, "manifest.json")))
TABLES = np.load(os.path.join(HERE, "golden", "lut_tables_k.npz"))
STILLS = sorted(n for n, e in MAN.items() if not e["frames"] and e["kind"] != "rgb")


def _coefficients(e):
    img = synth.noise_frame(e["W"], e["H"]) if e["kind"] == "noise" else synth.synth_frame(e["W"], e["H"], 0)
    px = O.pad_mirror(img)
    return O.truncate(O.dwt97_forward(px, e["wl"], e["qs"])) if e["type"] else O.dwt53_forward(px, e["wl"])


@pytest.mark.parametrize("name", STILLS)
def test_bulk_mode_encoder_matches_reference_binary(name):
    e = MAN[name]
    assert e["rc_enc"] == 0
    key = f"{e['lut']}_wl{e['wl']}"
    lut = O.Lut(TABLES[key + "_hdr"], TABLES[key])
    q = _coefficients(e)
    Ha, Wa = q.shape
    ncb = (Wa // 64) * (Ha // 64)
    cs, sizes = O.bpc_encode(q, e["wl"], lut, k=float(e["k"]))
    ref = np.fromfile(os.path.join(KDIR, name + ".enc"), dtype=np.uint16)
    straddles = (Wa >> e["wl"]) < 64 or (Ha >> e["wl"]) < 64
    if not straddles:
        mine = O.bsb_pack(cs, sizes, O.header_pack(e["W"], e["H"], e["wl"], e["type"], e["qs"], 0, k=float(e["k"])))
        assert mine.size == ref.size and np.array_equal(mine, ref)
        assert e["rc_dec"] == 0  # and the reference decodes its own stream
        return
    assert ref[8] == int(e["k"] * 1000)  # header word 8 carries k (BitStreamBuilder.cpp:82-93)
    rcs, rsizes, _ = O.bsb_unpack(ref, ncb)
    for b in range(1, ncb):  # block 0 straddles LL / HL of the deepest level
        assert sizes[b] == rsizes[b], (name, b)
        assert np.array_equal(cs[b * 4096: b * 4096 + sizes[b]], rcs[b * 4096: b * 4096 + rsizes[b]]), (name, b)


def test_reference_decoder_only_survives_uniform_blocks():
    """What the reference itself did with these streams (recorded, not recomputed): decode fails with an illegal
    memory access exactly where a codeblock straddles subbands and k leaves planes to the bulk pass."""
    for name, e in MAN.items():
        Wa, Ha = -(-e["W"] // 64) * 64, -(-e["H"] // 64) * 64
        straddles = (Wa >> e["wl"]) < 64 or (Ha >> e["wl"]) < 64
        if straddles and e["k"] < 65:
            assert e["rc_dec"] == 255 and "illegal memory access" in e["stdout_dec"], name
        else:
            assert e["rc_dec"] == 0 and (e["type"] or e["dec_equal_input"]), name


def _lengths(e):
    return [int(x) for x in e["size_text"].replace(",", " ").split()]


def test_bulk_mode_video_frames_match_reference_binary():
    """-video 1 -k 1: every frame is coded with the same rules, only frame 0 carries the header (k in word 8)."""
    name = "k1_v256_f3_wl2_q1"
    e = MAN[name]
    lut = O.Lut(TABLES[f"{e['lut']}_wl{e['wl']}_hdr"], TABLES[f"{e['lut']}_wl{e['wl']}"])
    ref = np.fromfile(os.path.join(KDIR, name + ".enc"), dtype=np.uint16)
    off = 0
    for f, n in enumerate(_lengths(e)):
        mine = O.encode_frame(synth.synth_frame(e["W"], e["H"], f), e["wl"], True, e["qs"], lut, first=(f == 0),
                              frames=e["frames"], k=float(e["k"]))
        assert mine.size == n and np.array_equal(mine, ref[off:off + n]), f
        off += n
    assert off == ref.size


def test_bulk_mode_rgb_components_match_reference_binary():
    """-isRGB 1 -k 1: component c uses the LUT sets of file letter R / G / B."""
    name = "k1_rgb256_wl2_ll"
    e = MAN[name]
    key = f"{e['lut']}_wl{e['wl']}"
    luts = [O.Lut(TABLES[key + "_hdr"], TABLES[key if c == 0 else f"{key}_c{c}"]) for c in range(3)]
    streams = O.encode_rgb_frame(synth.synth_rgb(e["W"], e["H"], 0), e["wl"], False, 1.0, luts, [True, False, False], 0,
                                 k=float(e["k"]))
    ref = np.fromfile(os.path.join(KDIR, name + ".enc"), dtype=np.uint16)
    off = 0
    for c, n in enumerate(_lengths(e)):
        assert streams[c].size == n and np.array_equal(streams[c], ref[off:off + n]), c
        off += n
    assert off == ref.size and e["rc_dec"] == 0 and e["dec_equal_input"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/CUDA_ImCod/LUT"), reason="reference tree not present")
def test_k_tables_fixture_matches_reference_files():
    for key in [k for k in TABLES.files if not k.endswith("_hdr")]:
        folder, rest = key.rsplit("_wl", 1)
        wl, comp = (rest.split("_c") + ["0"])[:2]
        lut = O.lut_load_k(f"/root/reference/CUDA_ImCod/LUT/{folder}/", int(wl), int(comp))
        assert np.array_equal(lut.table, TABLES[key])
        if comp == "0":
            assert np.array_equal(lut.hdr, TABLES[key + "_hdr"])


@pytest.mark.parametrize("name", STILLS)
def test_bulk_mode_decoder_inverts_the_encoder_and_matches_the_reference_decoder(name):
    """pcso_bpc_decode_k (decodeBulkProcessing, BPCEngine.cu:1487-1530,1776-1837): exact coefficient round trip for
    every case; for the geometries the reference can decode, the pixels decoded from the REFERENCE's stream are the
    ones the reference decoder wrote (lossless: the input; 9/7: the stored PGM)."""
    e = MAN[name]
    key = f"{e['lut']}_wl{e['wl']}"
    lut = O.Lut(TABLES[key + "_hdr"], TABLES[key])
    q = _coefficients(e)
    Ha, Wa = q.shape
    k = float(e["k"])
    cs, sizes = O.bpc_encode(q, e["wl"], lut, k=k)
    assert np.array_equal(O.bpc_decode(cs, sizes, Wa, Ha, e["wl"], lut, k=k), q)
    if e["rc_dec"] != 0:
        return
    ref = np.fromfile(os.path.join(KDIR, name + ".enc"), dtype=np.uint16)
    out = O.decode_frame(ref, Wa, Ha, e["wl"], bool(e["type"]), e["qs"], lut, k=k)[: e["H"], : e["W"]]
    if not e["type"]:
        img = synth.noise_frame(e["W"], e["H"]) if e["kind"] == "noise" else synth.synth_frame(e["W"], e["H"], 0)
        assert e["dec_equal_input"] and np.array_equal(out, img)
    else:
        d = np.fromfile(os.path.join(KDIR, name + ".dec"), dtype=np.uint8)
        assert np.array_equal(out, d[d.size - e["W"] * e["H"]:].reshape(e["H"], e["W"]))


def test_text_loaders_rebuild_k_tables(tmp_path):
    """The folder written back by lut_fixture.materialize_k parses to the same tables with the oracle's loader and
    with the product's (pcs_lut_load_k, C ABI) -- no reference tree needed."""
    import ctypes as C

    import lut_fixture
    from picsong_b200 import _lib
    lib = _lib.load()
    for folder, wl, comps in (("n1_lossless", 2, 3), ("n1_lossy", 2, 1), ("video_lossy", 2, 1)):
        d = lut_fixture.materialize_k(folder, wl, str(tmp_path))
        for c in range(comps):
            want = TABLES[f"{folder}_wl{wl}" if c == 0 else f"{folder}_wl{wl}_c{c}"]
            assert np.array_equal(O.lut_load_k(d, wl, c).table, want)
            hdr = (C.c_int32 * 8)()
            tab = (C.c_int32 * (1 << 20))()
            n = lib.pcs_lut_load_k(d.encode(), wl, c, hdr, tab, 1 << 20)
            assert n == want.size and np.array_equal(np.array(tab[:n], dtype=np.int32), want)
