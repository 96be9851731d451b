#!/usr/bin/env python
"""Test infrastructure.  Runs HERE (build container, /root/reference present) after
oracle/run_ref_golden.py has run on a B200: turns the reference binary's outputs
(gpurun_out/ref_golden/) and the reference's LUT text tables into committed fixtures:

  tests/golden/ref_streams.npz   codestreams (+ decoded payloads for lossy cases) the UNMODIFIED
                                 reference produced on a B200 for small synthetic inputs
  tests/golden/ref_manifest.json per-case parameters, sha256/length of every output incl. the
                                 BASELINE.json-sized cases (hash only), reference stdout tails
  tests/golden/lut_tables.npz    the k=0 greyscale ("R") probability tables of the four LUT folders,
                                 as parsed rows (lvl, sb, bp, values...) + header.txt values;
                                 tests/lut_fixture.py re-materialises the text folder from them
"""
import json, os, re, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "ref_golden")
DST = os.path.join(ROOT, "tests", "golden")
REF_LUT = "/root/reference/CUDA_ImCod/LUT"


def parse_rows(path):
    rows = []
    for line in open(path):
        line = line.strip()
        if not line:
            continue
        left, right = line.split(":")
        rows.append([int(t) for t in left.split()] + [int(t) for t in right.split()])
    return np.asarray(rows, dtype=np.int16)


def main():
    os.makedirs(DST, exist_ok=True)
    # Later GPU runs only re-run a subset of cases (and the video cases need the reference built with
    # its default host optimisation level and >= 2 streams), so manifests are merged: start from what is
    # already committed, then overlay every manifest found under gpurun_out/.
    man, where = {}, {}
    committed = os.path.join(DST, "ref_manifest.json")
    if os.path.exists(committed):
        man.update(json.load(open(committed)))
    for mpath, fdir in ((os.path.join(SRC, "manifest.json"), SRC),
                        (os.path.join(ROOT, "gpurun_out", "manifest_video_ns2.json"), os.path.join(ROOT, "gpurun_out", "video_ns2"))):
        if os.path.exists(mpath):
            for k, v in json.load(open(mpath)).items():
                if v.get("enc_len") or k not in man:
                    man[k] = v
                    where[k] = fdir
    arrays = {}
    old = os.path.join(DST, "ref_streams.npz")
    if os.path.exists(old):
        with np.load(old) as z:
            arrays.update({k: z[k] for k in z.files})
    out_man = {}
    for name, e in man.items():
        e = dict(e)
        e["stdout_enc"] = e["stdout_enc"][-400:]
        e["stdout_dec"] = e["stdout_dec"][-400:]
        fdir = where.get(name, SRC)
        enc = os.path.join(fdir, name + ".enc")
        if os.path.exists(enc) and os.path.getsize(enc) > 0 and os.path.getsize(enc) < (1 << 20):
            arrays[name + ".enc"] = np.fromfile(enc, dtype=np.uint16)
            dec = os.path.join(fdir, name + ".dec")
            npx = e["W"] * e["H"] * max(e["frames"], 1) * (3 if e.get("kind") == "rgb" else 1)
            if os.path.exists(dec) and (e["type"] == 1 or not e.get("dec_equal_input")):
                d = np.fromfile(dec, dtype=np.uint8)
                arrays[name + ".dec"] = d[d.size - npx:]
                e["dec_pgm_header"] = bytes(d[: d.size - npx]).decode("ascii")
        out_man[name] = e
    np.savez_compressed(os.path.join(DST, "ref_streams.npz"), **arrays)
    json.dump(out_man, open(os.path.join(DST, "ref_manifest.json"), "w"), indent=1, sort_keys=True)

    lut = {}
    for folder in ("n1_lossless", "n1_lossy", "video_lossless", "video_lossy"):
        hdr = [int(l.split(";")[1]) for l in open(os.path.join(REF_LUT, folder, "header.txt")) if ";" in l]
        lut[folder + ".header"] = np.asarray(hdr, dtype=np.int32)
        for sec in ("ref", "sig", "sign"):
            lut[f"{folder}.{sec}"] = parse_rows(os.path.join(REF_LUT, folder, f"{sec}R.txt_0"))
            for comp in "GB":  # RGB path: one table set per component (Engines/Engine.cu:124-139)
                lut[f"{folder}.{sec}{comp}"] = parse_rows(os.path.join(REF_LUT, folder, f"{sec}{comp}.txt_0"))
    np.savez_compressed(os.path.join(DST, "lut_tables.npz"), **lut)
    for f in sorted(os.listdir(DST)):
        print(f, os.path.getsize(os.path.join(DST, f)))


if __name__ == "__main__":
    main()
