"""Access to tests/golden: outputs of the UNMODIFIED reference binary run on a B200
(oracle/run_ref_golden.py -> oracle/make_golden.py)."""
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class Golden:
    def __init__(self):
        self.manifest = json.load(open(os.path.join(_HERE, "golden", "ref_manifest.json")))
        self._z = np.load(os.path.join(_HERE, "golden", "ref_streams.npz"))

    def stream(self, name):
        return self._z[name + ".enc"]

    def decoded(self, name):
        k = name + ".dec"
        return self._z[k] if k in self._z.files else None

    def names_with_streams(self, rgb=False):
        """Greyscale cases by default; rgb=True: the RGB-path cases (tests/test_gpu_rgb.py)."""
        return sorted(k[:-4] for k in self._z.files
                      if k.endswith(".enc") and (self.manifest[k[:-4]].get("kind") == "rgb") == rgb)

    def input(self, name):
        from picsong_b200 import synth
        e = self.manifest[name]
        fn = {"synth": synth.synth_frame, "zeros": synth.zeros_frame, "noise": synth.noise_frame}[e["kind"]]
        return fn(e["W"], e["H"])
