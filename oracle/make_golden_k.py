#!/usr/bin/env python
"""Test infrastructure -- NOT product code.  Runs in the build container (reads /root/reference/CUDA_ImCod/LUT).

Writes tests/golden/lut_tables_k.npz: for the k-factor reference runs of tests/golden/k_factor (made on the GPU box by
`oracle/run_ref_golden.py --k`), the probability tables of ALL bit-plane files of the LUT folder, parsed by the
oracle's loader (oracle.lut_load_k), so that the CPU test does not need the reference tree at run time."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

LUT = "/root/reference/CUDA_ImCod/LUT"
man = json.load(open(os.path.join(ROOT, "tests", "golden", "k_factor", "manifest.json")))
out = {}
for e in man.values():
    key = f"{e['lut']}_wl{e['wl']}"
    if key in out:
        continue
    lut = O.lut_load_k(f"{LUT}/{e['lut']}/", e["wl"])
    out[key] = lut.table
    out[key + "_hdr"] = lut.hdr
for e in man.values():  # RGB cases: component 1, 2 tables (files G, B) as well
    if e["kind"] != "rgb":
        continue
    for c in (1, 2):
        out[f"{e['lut']}_wl{e['wl']}_c{c}"] = O.lut_load_k(f"{LUT}/{e['lut']}/", e["wl"], c).table
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lut_tables_k.npz"), **out)
print({k: v.shape for k, v in out.items()})
