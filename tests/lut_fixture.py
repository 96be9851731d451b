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
