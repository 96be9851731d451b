"""GPU test of the PICSONG-compatible command line (C++ host code over the C ABI): same flags and file
formats as the reference's Launcher; outputs compared with the reference binary's (tests/golden)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cuda-image-and-video-codec_b200", "PICSONG_b200")


def run(args):
    p = subprocess.run([CLI] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    return p.returncode, p.stdout


def test_cli_still_and_video(tmp_path, golden, lut_root):
    from picsong_b200 import synth
    assert os.path.exists(CLI), "build the CLI with make -C cuda-image-and-video-codec_b200"
    # ---- still, lossless: byte-identical to the reference's file; decode gives the reference's PGM ----
    e = golden.manifest["s256x192_wl3_ll"]
    raw, enc, dec = tmp_path / "a.raw", tmp_path / "a.pcs", tmp_path / "a.pgm"
    golden.input("s256x192_wl3_ll").tofile(raw)
    rc, out = run(["-cd", 0, "-wl", 3, "-cp", 2, "-type", 0, "-qs", 1, "-i", raw, "-o", enc, "-cbWidth", 64, "-cbLength", 18,
                   "-xSize", 256, "-ySize", 192, "-video", 0, "-LUTFolder", f"{lut_root}/n1_lossless/", "-k", 0])
    assert rc == 0, out
    assert "User entered -wl command 3" in out and "The time spent with the app is:" in out
    assert np.array_equal(np.fromfile(enc, dtype=np.uint16), golden.stream("s256x192_wl3_ll"))
    assert not os.path.exists(str(enc) + "_SIZE")
    rc, out = run(["-cd", 1, "-i", enc, "-o", dec, "-LUTFolder", f"{lut_root}/n1_lossless/", "-video", 0])
    assert rc == 0, out
    blob = open(dec, "rb").read()
    assert blob.startswith(b"P5\n256 192\n255\n")
    assert np.array_equal(np.frombuffer(blob[len(b"P5\n256 192\n255\n"):], dtype=np.uint8), golden.input("s256x192_wl3_ll").reshape(-1))
    # ---- video, lossy: stream + _SIZE sidecar identical to the reference's ----
    e = golden.manifest["v256_f3_wl3_q1"]
    raw, enc, dec = tmp_path / "v.raw", tmp_path / "v.pcs", tmp_path / "v.out"
    synth.synth_video(256, 256, 3).tofile(raw)
    rc, out = run(["-cd", 0, "-wl", 3, "-cp", 2, "-type", 1, "-qs", 1.0, "-i", raw, "-o", enc, "-cbWidth", 64, "-cbHeight", 18,
                   "-xSize", 256, "-ySize", 256, "-video", 1, "-frames", 3, "-numberOfStreams", 2, "-LUTFolder", f"{lut_root}/video_lossy/"])
    assert rc == 0, out
    assert np.array_equal(np.fromfile(enc, dtype=np.uint16), golden.stream("v256_f3_wl3_q1"))
    assert open(str(enc) + "_SIZE").read() == e["size_text"]
    rc, out = run(["-cd", 1, "-i", enc, "-o", dec, "-LUTFolder", f"{lut_root}/video_lossy/", "-video", 1])
    assert rc == 0, out
    d = np.abs(np.fromfile(dec, dtype=np.uint8).astype(int) - golden.decoded("v256_f3_wl3_q1").astype(int))
    assert d.max() <= 1 and (d != 0).mean() <= 1e-4
    # ---- invalid parameters: the reference's message and exit(-1) ----
    rc, out = run(["-cd", 0, "-wl", 11, "-i", raw, "-o", enc, "-xSize", 256, "-ySize", 256])
    assert rc == 255 and "Incorrect parameters. Please choose valid values." in out


def test_cli_rgb(tmp_path, golden, lut_root):
    """-isRGB 1: planar RGB in, stream + _SIZE identical to the reference binary's, planar RAW out (README.md:105-115)."""
    from picsong_b200 import synth
    e = golden.manifest["rgb256x192_wl3_ll"]
    raw, enc, dec = tmp_path / "c.raw", tmp_path / "c.pcs", tmp_path / "c.out"
    synth.synth_rgb(256, 192).tofile(raw)
    rc, out = run(["-cd", 0, "-wl", 3, "-cp", 2, "-type", 0, "-qs", 1, "-i", raw, "-o", enc, "-cbWidth", 64, "-cbHeight", 18,
                   "-xSize", 256, "-ySize", 192, "-video", 0, "-isRGB", 1, "-LUTFolder", f"{lut_root}/n1_lossless/", "-k", 0])
    assert rc == 0, out
    assert np.array_equal(np.fromfile(enc, dtype=np.uint16), golden.stream("rgb256x192_wl3_ll"))
    assert open(str(enc) + "_SIZE").read() == e["size_text"]
    rc, out = run(["-cd", 1, "-i", enc, "-o", dec, "-LUTFolder", f"{lut_root}/n1_lossless/", "-video", 0])
    assert rc == 0, out
    assert np.array_equal(np.fromfile(dec, dtype=np.uint8), synth.synth_rgb(256, 192).reshape(-1))
    # RGB video, lossy
    e = golden.manifest["rgbv256_f2_wl3_q1"]
    raw, enc, dec = tmp_path / "cv.raw", tmp_path / "cv.pcs", tmp_path / "cv.out"
    np.stack([synth.synth_rgb(256, 256, f) for f in range(2)]).tofile(raw)
    rc, out = run(["-cd", 0, "-wl", 3, "-cp", 2, "-type", 1, "-qs", 1.0, "-i", raw, "-o", enc, "-cbWidth", 64, "-cbHeight", 18,
                   "-xSize", 256, "-ySize", 256, "-video", 1, "-frames", 2, "-isRGB", 1, "-numberOfStreams", 2,
                   "-LUTFolder", f"{lut_root}/video_lossy/"])
    assert rc == 0, out
    assert np.array_equal(np.fromfile(enc, dtype=np.uint16), golden.stream("rgbv256_f2_wl3_q1"))
    assert open(str(enc) + "_SIZE").read() == e["size_text"]
    rc, out = run(["-cd", 1, "-i", enc, "-o", dec, "-LUTFolder", f"{lut_root}/video_lossy/", "-video", 1])
    assert rc == 0, out
    d = np.abs(np.fromfile(dec, dtype=np.uint8).astype(int) - golden.decoded("rgbv256_f2_wl3_q1").astype(int))
    assert d.max() <= 1 and (d != 0).mean() <= 1e-4


def test_cli_k_factor(tmp_path):
    """-k 1: the complexity-scalable mode through the command line; stream identical to the reference binary's
    (tests/golden/k_factor), the decoder takes k from the header and reproduces the input."""
    import json

    import lut_fixture
    from picsong_b200 import synth
    kdir = os.path.join(ROOT, "tests", "golden", "k_factor")
    e = json.load(open(os.path.join(kdir, "manifest.json")))["k1_s256_wl2_ll"]
    lut = lut_fixture.materialize_k("n1_lossless", 2, str(tmp_path))
    raw, enc, dec = tmp_path / "k.raw", tmp_path / "k.pcs", tmp_path / "k.pgm"
    img = synth.synth_frame(256, 256, 0)
    img.tofile(raw)
    rc, out = run(["-cd", 0, "-wl", 2, "-cp", 2, "-type", 0, "-qs", 1, "-i", raw, "-o", enc, "-cbWidth", 64, "-cbHeight", 18,
                   "-xSize", 256, "-ySize", 256, "-video", 0, "-LUTFolder", lut, "-k", 1])
    assert rc == 0, out
    assert np.array_equal(np.fromfile(enc, dtype=np.uint16), np.fromfile(os.path.join(kdir, "k1_s256_wl2_ll.enc"), dtype=np.uint16))
    rc, out = run(["-cd", 1, "-i", enc, "-o", dec, "-LUTFolder", lut, "-video", 0])
    assert rc == 0, out
    blob = open(dec, "rb").read()
    assert np.array_equal(np.frombuffer(blob[len(b"P5\n256 256\n255\n"):], dtype=np.uint8), img.reshape(-1)) and e["rc_dec"] == 0


def test_cli_many_batches_wrap_the_ring(tmp_path, lut_root):
    """The command line's reader / GPU / writer pipeline circulates three page-locked buffers; PICSONG_BATCH forces small
    batches so that a 13-frame video goes through 7 (encode) and 5 (decode) batches and every buffer is reused.  The
    file must equal what ONE pcs_encode_frames call over all frames produces, and decode must return the input."""
    from picsong_b200 import codec, synth
    W, H, wl, n = 256, 192, 3, 13
    frames = synth.synth_video(W, H, n)
    raw, enc, dec = tmp_path / "m.raw", tmp_path / "m.pcs", tmp_path / "m.out"
    frames.tofile(raw)
    args = ["-cd", 0, "-wl", wl, "-cp", 2, "-type", 0, "-qs", 1, "-i", raw, "-o", enc, "-cbWidth", 64, "-cbHeight", 18, "-xSize", W, "-ySize", H,
            "-video", 1, "-frames", n, "-LUTFolder", f"{lut_root}/video_lossless/"]
    p = subprocess.run([CLI] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300,
                       env=dict(os.environ, PICSONG_BATCH="2"))
    assert p.returncode == 0, p.stdout
    assert "Start-up (CUDA context" in p.stdout
    cd = codec.Codec(codec.Params(W, H, wl, False, 1.0, frames=n, max_batch=16), codec.load_lut(f"{lut_root}/video_lossless/", wl))
    s, lens = cd.encode_frames(frames)
    cd.close()
    assert np.array_equal(np.fromfile(enc, dtype=np.uint16), s)
    assert open(str(enc) + "_SIZE").read() == ",".join(str(int(v)) for v in lens)
    p = subprocess.run([CLI, "-cd", "1", "-i", str(enc), "-o", str(dec), "-LUTFolder", f"{lut_root}/video_lossless/", "-video", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300, env=dict(os.environ, PICSONG_BATCH="3"))
    assert p.returncode == 0, p.stdout
    assert np.array_equal(np.fromfile(dec, dtype=np.uint8).reshape(n, H, W), frames)


def test_cli_rgb_still_with_frames_flag(tmp_path, golden, lut_root):
    """`-video 0 -isRGB 1 -frames N`: the reference codes the components of a still with iter = component, so only plane 0
    carries the header whatever -frames says.  The stream must equal the plain RGB still's except for the
    header's frame field."""
    from picsong_b200 import synth
    raw, enc0, enc5 = tmp_path / "c.raw", tmp_path / "c0.pcs", tmp_path / "c5.pcs"
    synth.synth_rgb(256, 192).tofile(raw)
    base = ["-cd", 0, "-wl", 3, "-cp", 2, "-type", 0, "-qs", 1, "-i", raw, "-cbWidth", 64, "-cbHeight", 18, "-xSize", 256, "-ySize", 192,
            "-video", 0, "-isRGB", 1, "-LUTFolder", f"{lut_root}/n1_lossless/", "-k", 0]
    rc, out = run(base + ["-o", enc0]); assert rc == 0, out
    rc, out = run(base + ["-o", enc5, "-frames", 5]); assert rc == 0, out
    a, b = np.fromfile(enc0, dtype=np.uint16), np.fromfile(enc5, dtype=np.uint16)
    assert a.size == b.size and np.array_equal(a[9:], b[9:])          # planes 1 and 2 start with 0xFFFF x 9 in both
    assert open(str(enc0) + "_SIZE").read() == open(str(enc5) + "_SIZE").read()
    assert np.array_equal(a, golden.stream("rgb256x192_wl3_ll"))
