#!/usr/bin/env python
"""Test infrastructure -- NOT product code.  Runs ON THE GPU BOX (via gpurun).

Drives the unmodified reference binary (oracle/_ref/PICSONG_ref, built by oracle/build_ref.sh)
over the synthetic inputs of SURVEY.md section 8d and collects what it produces:
codestream files, `_SIZE` sidecars, decoded PGM/RAW, stdout, exit codes.  Small cases are kept
whole (they become tests/golden fixtures via oracle/make_golden.py); big cases keep sha256 +
length only.  Output: gpurun_out/ref_golden/.
"""
import hashlib, json, os, subprocess, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-image-and-video-codec_b200"))
from picsong_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "PICSONG_ref")
LUT = os.path.join(ROOT, "oracle", "_ref", "LUT")
OUT = os.path.join(ROOT, "gpurun_out", "ref_golden")
TMP = "/tmp/ref_golden"

# name, kind, W, H, frames, wl, type, qs, lut, keep_files
CASES = [
    ("s128_wl2_ll",   "synth", 128, 128, 0, 2, 0, 1.0, "n1_lossless", True),
    ("s256x192_wl3_ll", "synth", 256, 192, 0, 3, 0, 1.0, "n1_lossless", True),
    ("s512_wl4_ll",   "synth", 512, 512, 0, 4, 0, 1.0, "n1_lossless", True),
    ("s512_wl5_ll",   "synth", 512, 512, 0, 5, 0, 1.0, "n1_lossless", True),
    ("s1024_wl5_ll",  "synth", 1024, 1024, 0, 5, 0, 1.0, "n1_lossless", False),
    ("pad200x150_wl2_ll", "synth", 200, 150, 0, 2, 0, 1.0, "n1_lossless", True),
    ("zero128_wl2_ll", "zeros", 128, 128, 0, 2, 0, 1.0, "n1_lossless", True),
    ("noise128_wl2_ll", "noise", 128, 128, 0, 2, 0, 1.0, "n1_lossless", True),
    ("noise256_wl1_ll", "noise", 256, 256, 0, 1, 0, 1.0, "n1_lossless", True),
    ("s256_wl3_q1",   "synth", 256, 256, 0, 3, 1, 1.0, "n1_lossy", True),
    ("s256_wl3_q05",  "synth", 256, 256, 0, 3, 1, 0.5, "n1_lossy", True),
    ("s512_wl4_q01",  "synth", 512, 512, 0, 4, 1, 0.1, "n1_lossy", True),
    ("noise128_wl2_q1", "noise", 128, 128, 0, 2, 1, 1.0, "n1_lossy", True),
    ("v128_f4_wl2_ll", "synth", 128, 128, 4, 2, 0, 1.0, "video_lossless", True),
    ("v256_f3_wl3_q1", "synth", 256, 256, 3, 3, 1, 1.0, "video_lossy", True),
    # BASELINE.json configs (hash only)
    ("cfg2_4096_wl5_q1",  "synth", 4096, 4096, 0, 5, 1, 1.0, "n1_lossy", False),
    ("cfg2_4096_wl5_q05", "synth", 4096, 4096, 0, 5, 1, 0.5, "n1_lossy", False),
    ("cfg2_4096_wl5_q01", "synth", 4096, 4096, 0, 5, 1, 0.1, "n1_lossy", False),
    ("cfg2_4096_wl5_ll",  "synth", 4096, 4096, 0, 5, 0, 1.0, "n1_lossless", False),
    ("cfg3_8k_wl6_ll",    "synth", 7680, 4320, 0, 6, 0, 1.0, "n1_lossless", False),
    ("cfg4_4k_f4_wl5_q1", "synth", 3840, 2160, 4, 5, 1, 1.0, "video_lossy", False),
    ("cfg5_band_f2_wl5_ll", "band", 32768, 4096, 2, 5, 0, 1.0, "n1_lossless", False),
    # RGB path: planar R,G,B planes per frame, `-isRGB 1`, per-component LUT files
    ("rgb128_wl2_ll", "rgb", 128, 128, 0, 2, 0, 1.0, "n1_lossless", True),
    ("rgb256x192_wl3_ll", "rgb", 256, 192, 0, 3, 0, 1.0, "n1_lossless", True),
    ("rgb256x192_wl3_q1", "rgb", 256, 192, 0, 3, 1, 1.0, "n1_lossy", True),
    ("rgb200x150_wl2_q05", "rgb", 200, 150, 0, 2, 1, 0.5, "n1_lossy", True),
    ("rgbv128_f3_wl2_ll", "rgb", 128, 128, 3, 2, 0, 1.0, "video_lossless", True),
    ("rgbv256_f2_wl3_q1", "rgb", 256, 256, 2, 3, 1, 1.0, "video_lossy", True),
    ("rgb2048_wl5_q1", "rgb", 2048, 2048, 0, 5, 1, 1.0, "n1_lossy", False),
    ("rgb2048_wl5_ll", "rgb", 2048, 2048, 0, 5, 0, 1.0, "n1_lossless", False),
]


# k-factor ("bulk") mode, SURVEY 8f.2: not implemented by the new code yet; these runs pin the reference's behaviour
# for the oracle restatement that has to come first.  Run with `--k`; needs oracle/_ref/LUT_all (build_ref.sh).
# name, kind, W, H, frames, wl, type, qs, lut, keep_files, k
K_CASES = [
    ("k05_s128_wl2_ll", "synth", 128, 128, 0, 2, 0, 1.0, "n1_lossless", True, 0.5),
    ("k1_s128_wl2_ll", "synth", 128, 128, 0, 2, 0, 1.0, "n1_lossless", True, 1.0),
    ("k3_s128_wl2_ll", "synth", 128, 128, 0, 2, 0, 1.0, "n1_lossless", True, 3.0),
    ("k2_noise128_wl2_ll", "noise", 128, 128, 0, 2, 0, 1.0, "n1_lossless", True, 2.0),
    ("k1_s256_wl3_q1", "synth", 256, 256, 0, 3, 1, 1.0, "n1_lossy", True, 1.0),
    ("k07_s256x192_wl3_ll", "synth", 256, 192, 0, 3, 0, 1.0, "n1_lossless", True, 0.7),
    ("k1_v128_f3_wl2_q1", "synth", 128, 128, 3, 2, 1, 1.0, "video_lossy", True, 1.0),
    ("k65_s128_wl2_ll", "synth", 128, 128, 0, 2, 0, 1.0, "n1_lossless", True, 65.0),
    # no codeblock straddles a subband here (LL2 is 64x64): every lane of a warp leaves the plane loop at the same plane
    ("k05_s256_wl2_ll", "synth", 256, 256, 0, 2, 0, 1.0, "n1_lossless", True, 0.5),
    ("k1_s256_wl2_ll", "synth", 256, 256, 0, 2, 0, 1.0, "n1_lossless", True, 1.0),
    ("k2_s256_wl2_ll", "synth", 256, 256, 0, 2, 0, 1.0, "n1_lossless", True, 2.0),
    ("k1_s256_wl2_q1", "synth", 256, 256, 0, 2, 1, 1.0, "n1_lossy", True, 1.0),
    ("k1_noise256_wl2_ll", "noise", 256, 256, 0, 2, 0, 1.0, "n1_lossless", True, 1.0),
    ("k0_s256_wl2_ll", "synth", 256, 256, 0, 2, 0, 1.0, "n1_lossless", True, 0),
    ("k1_v256_f3_wl2_q1", "synth", 256, 256, 3, 2, 1, 1.0, "video_lossy", True, 1.0),
    ("k1_rgb256_wl2_ll", "rgb", 256, 256, 0, 2, 0, 1.0, "n1_lossless", True, 1.0),
]


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def make_input(kind, W, H, frames, path):
    n = max(frames, 1)
    with open(path, "wb") as f:
        for i in range(n):
            if kind == "synth":
                a = synth.synth_frame(W, H, i if frames else 0)
            elif kind == "band":
                a = synth.synth_frame(W, H, 0, 0, H * i)
            elif kind == "rgb":
                a = synth.synth_rgb(W, H, i if frames else 0)  # [3][H][W] planar
            elif kind == "zeros":
                a = synth.zeros_frame(W, H)
            else:
                a = synth.noise_frame(W, H)
            f.write(a.tobytes())


def run(cmd, timeout=600):
    t0 = time.time()
    try:
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout)
        return p.returncode, p.stdout.decode(errors="replace"), time.time() - t0
    except subprocess.TimeoutExpired as e:
        return -999, (e.stdout or b"").decode(errors="replace") + "\nTIMEOUT", time.time() - t0


def main():
    kmode = "--k" in sys.argv[1:]
    only = set(a for a in sys.argv[1:] if a != "--k")
    out_dir = OUT + ("_k" if kmode else "")
    lut_root = os.path.join(ROOT, "oracle", "_ref", "LUT_all") if kmode else LUT
    os.makedirs(out_dir, exist_ok=True)
    os.makedirs(TMP, exist_ok=True)
    manifest = {}
    for case in (K_CASES if kmode else CASES):
        name, kind, W, H, frames, wl, typ, qs, lut, keep = case[:10]
        kfac = case[10] if len(case) > 10 else 0
        if only and name not in only:
            continue
        raw = os.path.join(TMP, name + ".raw")
        enc = os.path.join(TMP, name + ".pcs")
        dec = os.path.join(TMP, name + (".dec.raw" if frames else ".dec.pgm"))
        for p in (enc, enc + "_SIZE", dec):
            if os.path.exists(p):
                os.remove(p)
        make_input(kind, W, H, frames, raw)
        video = 1 if frames else 0
        lutdir = os.path.join(lut_root, lut) + "/"
        ecmd = [REF, "-cd", "0", "-wl", str(wl), "-cp", "2", "-type", str(typ), "-qs", repr(qs),
                "-i", raw, "-o", enc, "-cbWidth", "64", "-cbHeight", "18", "-xSize", str(W),
                "-ySize", str(H), "-video", str(video), "-LUTFolder", lutdir, "-k", repr(kfac) if kfac else "0",
                "-numberOfStreams", os.environ.get("REF_STREAMS", "1")]
        if frames:
            ecmd += ["-frames", str(frames)]
        if kind == "rgb":
            ecmd += ["-isRGB", "1"]
        rc_e, out_e, t_e = run(ecmd)
        dcmd = [REF, "-cd", "1", "-i", enc, "-o", dec, "-LUTFolder", lutdir, "-video", str(video),
                "-numberOfStreams", os.environ.get("REF_STREAMS", "1")]
        rc_d, out_d, t_d = (None, "", 0.0)
        if os.path.exists(enc):
            rc_d, out_d, t_d = run(dcmd)
        ent = {"kind": kind, "W": W, "H": H, "frames": frames, "wl": wl, "type": typ, "qs": qs,
               "lut": lut, "k": kfac, "rc_enc": rc_e, "rc_dec": rc_d, "t_enc_s": round(t_e, 3),
               "t_dec_s": round(t_d, 3), "stdout_enc": out_e[-3000:], "stdout_dec": out_d[-3000:],
               "input_sha256": sha(raw)}
        for tag, p in (("enc", enc), ("size", enc + "_SIZE"), ("dec", dec)):
            if os.path.exists(p):
                ent[tag + "_len"] = os.path.getsize(p)
                ent[tag + "_sha256"] = sha(p)
                if tag == "size":
                    ent["size_text"] = open(p).read()
                if keep or os.environ.get("KEEP_ALL"):
                    os.replace(p, os.path.join(out_dir, name + "." + tag))
        if os.path.exists(dec) or (keep and os.path.exists(os.path.join(out_dir, name + ".dec"))):
            # decoded-vs-input comparison (payload only: PGM header is filesize - W*H*n bytes)
            pd = dec if os.path.exists(dec) else os.path.join(out_dir, name + ".dec")
            d = np.fromfile(pd, dtype=np.uint8)
            n = W * H * max(frames, 1) * (3 if kind == "rgb" else 1)
            x = np.fromfile(raw, dtype=np.uint8)
            if d.size >= n:
                d = d[d.size - n:] if not (frames or kind == "rgb") else d[:n]
                diff = d.astype(np.int32) - x.astype(np.int32)
                mse = float(np.mean(diff.astype(np.float64) ** 2))
                ent["dec_equal_input"] = bool(mse == 0.0)
                ent["dec_mse"] = mse
                ent["dec_psnr"] = (10 * np.log10(255.0 ** 2 / mse)) if mse > 0 else None
                ent["dec_maxabs"] = int(np.abs(diff).max())
        manifest[name] = ent
        print(name, "enc rc", rc_e, "dec rc", rc_d, {k: ent.get(k) for k in
              ("enc_len", "dec_equal_input", "dec_psnr", "t_enc_s", "t_dec_s")}, flush=True)
        for p in (raw, enc, enc + "_SIZE", dec):
            if os.path.exists(p):
                os.remove(p)
    with open(os.path.join(out_dir, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
