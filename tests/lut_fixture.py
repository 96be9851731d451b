"""Re-materialises the reference's LUT text folders from tests/golden/lut_tables.npz.

The fixture (made by oracle/make_golden.py from /root/reference/CUDA_ImCod/LUT) stores each table
file as parsed rows; this writes them back in the reference's on-disk format
(`header.txt` with eight `NAME;value` lines, `{ref,sig,sign}R.txt_0` with `lvl sb bp : v ... ` rows) so
that the text loaders (oracle and product) are exercised exactly as with the reference's folders.
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
FOLDERS = ("n1_lossless", "n1_lossy", "video_lossless", "video_lossy")
_HDR_NAMES = ("LUT_N_BITPLANES", "LUT_N_SUBBANDS", "N_CONTEXT_REFINEMENT", "N_CONTEXT_SIGN",
              "N_CONTEXT_SIGNIFICANCE", "MULT_PRECISION", "LUT_N_FILES", "AMOUNT_OF_BITPLANE_FILES")


def materialize(folder: str, dst_root: str) -> str:
    z = np.load(os.path.join(_HERE, "golden", "lut_tables.npz"))
    d = os.path.join(dst_root, folder)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "header.txt"), "w") as f:
        f.write("\n".join(f"{n};{int(v)}" for n, v in zip(_HDR_NAMES, z[folder + ".header"])))
    for sec in ("ref", "sig", "sign"):
        for comp in "RGB":
            key = f"{folder}.{sec}" + ("" if comp == "R" else comp)
            if key not in z.files:
                continue
            with open(os.path.join(d, f"{sec}{comp}.txt_0"), "w") as f:
                for r in z[key]:
                    f.write(f"{r[0]} {r[1]} {r[2]} : " + " ".join(str(int(v)) for v in r[3:]) + " \n")
    return d + "/"


def materialize_all(dst_root: str):
    return {f: materialize(f, dst_root) for f in FOLDERS}


def materialize_k(folder: str, wl: int, dst_root: str) -> str:
    """A LUT folder with ALL bit-plane files (`*_0 .. *_(n-1)`, the complexity-scalable mode) rebuilt from the parsed
    tables of tests/golden/lut_tables_k.npz (made by oracle/make_golden_k.py for exactly this folder and level count).
    Every (group, bit-plane) row is written out, so the text loaders re-create the same tables."""
    z = np.load(os.path.join(_HERE, "golden", "lut_tables_k.npz"))
    key = f"{folder}_wl{wl}"
    hdr = [int(v) for v in z[key + "_hdr"]]
    nBp, nSub, cR, cG, cS, files = hdr[0], hdr[1], hdr[2], hdr[3], hdr[4], hdr[7]
    groups = nSub * wl + 1
    R, S, G = groups * nBp * cR, groups * nBp * cS, groups * nBp * cG
    per = R + S + G
    d = os.path.join(dst_root, f"{folder}_k_wl{wl}")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "header.txt"), "w") as f:
        f.write("\n".join(f"{n};{v}" for n, v in zip(_HDR_NAMES, hdr)))
    for comp, letter in enumerate("RGB"):
        ck = key if comp == 0 else f"{key}_c{comp}"
        if ck not in z.files:
            continue
        tab = z[ck].reshape(files, per)
        for j in range(files):
            for sec, base, ctx in (("ref", 0, cR), ("sig", R, cS), ("sign", R + S, cG)):
                with open(os.path.join(d, f"{sec}{letter}.txt_{j}"), "w") as f:
                    for g in range(groups):
                        for bp in range(nBp):
                            v = tab[j, base + (g * nBp + bp) * ctx: base + (g * nBp + bp + 1) * ctx]
                            f.write(f"{g // nSub} {g % nSub} {bp} : " + " ".join(str(int(x)) for x in v) + " \n")
    return d + "/"
