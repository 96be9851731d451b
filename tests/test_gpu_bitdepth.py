"""GPU parity of -bps / -signedOrUnsigned: the CUDA path with a level shift of 2^(bps-1) (or 0 for signed
samples) against the CPU oracle and the reference executable's own runs (tests/golden/f4_ref.npz), through the C ABI and
through the command line."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import oracle as O  # noqa: E402  (checker only)
from picsong_b200 import codec  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cuda-image-and-video-codec_b200", "PICSONG_b200")
CASES = [("bps16_ll", 0, 16, 0), ("bps12_ll", 0, 12, 0), ("signed_ll", 0, 8, 1), ("signed_q1", 1, 8, 1), ("bps16_q1", 1, 16, 0)]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_level_shift_variants(case, lut_root):
    name, lossy, bps, signed = case
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "f4_ref.npz"))
    img = g["input"]
    H, W = img.shape
    folder = "n1_lossy" if lossy else "n1_lossless"
    olut, plut = O.lut_load(f"{lut_root}/{folder}/", 3), codec.load_lut(f"{lut_root}/{folder}/", 3)
    cd = codec.Codec(codec.Params(W, H, 3, bool(lossy), 1.0, extra={"bit_depth": bps, "bits_per_sample": bps, "is_signed": signed}), plut)
    stream, lens = cd.encode_frames(img[None])
    try:
        O.set_sample_offset(0 if signed else 1 << (bps - 1))
        expect = O.encode_frame_params(img, 3, bool(lossy), 1.0, olut, bitDepth=bps, bps=bps, isSigned=signed)
        assert np.array_equal(stream, expect), f"{name}: stream differs from the oracle"
        dec = cd.decode_frames(stream, lens)[0]
        assert np.array_equal(dec, O.decode_frame(expect, W, H, 3, bool(lossy), 1.0, olut)), f"{name}: pixels differ from the oracle"
    finally:
        O.set_sample_offset(128)
    ref = g[name + ".stream"]
    if name != "bps16_q1":
        assert np.array_equal(stream, ref), f"{name}: stream differs from the reference executable's"
        rdec = cd.decode_frames(ref, np.array([ref.size], dtype=np.int32))[0]
        assert np.array_equal(rdec, g[name + ".dec"]), name
    else:  # 9/7 with a 32768 shift: within the north_star tolerance of the reference (tests/golden/ref_runs/README.md)
        assert stream.size == ref.size and (stream != ref).mean() < 0.05
        psnr = lambda a: 10 * np.log10(255.0 ** 2 / np.mean((a.astype(np.float64) - img) ** 2))  # noqa: E731
        assert abs(psnr(dec) - psnr(g[name + ".dec"])) <= 0.01
    if not lossy:
        assert np.array_equal(dec, img)
    cd.close()


def test_cli_bps_and_signed(tmp_path, lut_root):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "f4_ref.npz"))
    img = g["input"]
    H, W = img.shape
    raw = tmp_path / "in.raw"
    img.tofile(raw)
    for name, flags in (("bps16_ll", ["-bps", "16"]), ("signed_ll", ["-signedOrUnsigned", "1"])):
        enc, dec = tmp_path / (name + ".pcs"), tmp_path / (name + ".pgm")
        p = subprocess.run([CLI, "-cd", "0", "-wl", "3", "-cp", "2", "-type", "0", "-qs", "1.0", "-i", str(raw), "-o", str(enc), "-cbWidth", "64",
                            "-cbHeight", "18", "-xSize", str(W), "-ySize", str(H), "-video", "0", "-LUTFolder", f"{lut_root}/n1_lossless/", "-k", "0"] + flags,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        assert p.returncode == 0, p.stdout
        assert np.array_equal(np.fromfile(enc, dtype=np.uint16), g[name + ".stream"]), name
        p = subprocess.run([CLI, "-cd", "1", "-i", str(enc), "-o", str(dec), "-LUTFolder", f"{lut_root}/n1_lossless/", "-video", "0"],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        assert p.returncode == 0, p.stdout
        blob = open(dec, "rb").read()
        hdr = bytes(g[name + ".pgm_header"])
        assert blob.startswith(hdr), (blob[:20], hdr)          # P5 / size / maxval = 2^bps - 1, as the reference writes it
        assert np.array_equal(np.frombuffer(blob[len(hdr):], dtype=np.uint8).reshape(H, W), img)
