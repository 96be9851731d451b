"""CPU tests of the C ABI boundary: the shared library loads, exports every symbol
include/picsong_b200.h declares, its host utilities agree with the oracle, and it fails loudly (no CPU
fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import oracle as O
from picsong_b200 import _lib, codec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "picsong_b200.h")).read()
    declared = set(re.findall(r"\b(pcs_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"pcs_ctx", "pcs_params"}
    lib = _lib.load()
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in picsong_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))


def test_host_utilities_match_oracle(lut_root):
    assert codec.adapted_dims(200, 150) == O.adapted_dims(200, 150) == (256, 192)
    assert codec.adapted_dims(7680, 4320) == (7680, 4352)
    for folder in ("n1_lossless", "n1_lossy", "video_lossless", "video_lossy"):
        for wl in (1, 2, 4, 5, 6):
            a = codec.load_lut(f"{lut_root}/{folder}/", wl)
            b = O.lut_load(f"{lut_root}/{folder}/", wl)
            assert np.array_equal(a.table, b.table) and np.array_equal(a.header, b.hdr)
    lut = codec.load_lut(f"{lut_root}/n1_lossy/", 5)
    for (W, H, wl, lossy, qs, frames) in ((512, 512, 5, 0, 1.0, 0), (4096, 4096, 5, 1, 0.5, 0), (3840, 2160, 5, 1, 1.0, 600),
                                          (7680, 4320, 6, 0, 1.0, 0), (200, 150, 2, 1, 0.1234, 3)):
        p = codec.Params(W, H, wl, bool(lossy), qs, frames=frames)
        h = codec.header_pack(p, lut)
        assert np.array_equal(h, O.header_pack(W, H, wl, lossy, qs, frames))
        u = codec.header_unpack(h)
        assert (u["width"], u["height"], u["wavelet_levels"], u["lossy"], u["frames"]) == (W, H, wl, lossy, frames)
        assert abs(u["quant_size"] - int(np.float32(qs) * 10000) / 10000.0) < 1e-7


def test_lut_errors(tmp_path):
    with pytest.raises(codec.PcsError):
        codec.load_lut(str(tmp_path) + "/nope/", 5)


def test_parameter_validation_and_no_cpu_fallback(lut_root):
    """pcs_create mirrors Launcher.cu:132 ("Incorrect parameters") and never falls back to the CPU."""
    lut = codec.load_lut(f"{lut_root}/n1_lossless/", 5)
    lib = _lib.load()

    def create(**kw):
        p = codec.Params(512, 512, 5)
        for k, v in kw.items():
            setattr(p, k, v)
        cp = p.to_c(lut)
        ctx = C.c_void_p()
        return lib.pcs_create(C.byref(cp), C.byref(ctx)), ctx

    assert create(wl=0)[0] == -1 and create(wl=11)[0] == -1        # PCS_E_ARG
    assert create(cb_height=17)[0] == -1 and create(cb_width=65)[0] == -1
    assert create(qs=1.5)[0] == -1 and create(width=0)[0] == -1
    assert create(cp=3)[0] == -4                                    # PCS_E_UNSUPPORTED (out of scope)
    assert create(k=1.0)[0] == -2                                   # PCS_E_LUT: k > 0 needs the tables of all bit-plane files
    import torch
    if not torch.cuda.is_available():
        rc, ctx = create()
        assert rc > 0 and not ctx.value, "without a CUDA device pcs_create must fail with the CUDA error"


def test_torch_frame_generator_matches_numpy():
    """bench.py generates the 600-frame video of BASELINE configs[3] with torch integer ops (on the GPU there); the formula is
    integer-only, so it must reproduce picsong_b200.synth bit for bit -- also through bench.gen_frames' index runs and the
    row bands of configs[4]."""
    import os
    import sys

    import numpy as np
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from picsong_b200 import synth
    assert np.array_equal(synth.synth_video_torch(200, 150, 3, 7).numpy(), synth.synth_video(200, 150, 3, 7))
    cfg = {"W": 192, "H": 130}
    idx = [4, 5, 6, 20, 21, 9]
    got = bench.gen_frames(cfg, idx, torch.device("cpu")).numpy()
    assert np.array_equal(got, np.stack([synth.synth_frame(192, 130, f) for f in idx]))
    bands = bench.gen_frames({"W": 256, "H": 64, "bands": True}, [0, 3], torch.device("cpu")).numpy()
    assert np.array_equal(bands, np.stack([synth.synth_frame(256, 64, 0, 0, 64 * b) for b in (0, 3)]))


def test_create_refuses_tables_with_more_rows_than_the_decoder_stages():
    """the decoder stages at most kMaxProbRows probability rows in shared memory while the encoder read
    every row from global memory, so a custom LUT with more rows (nBp = 16 at wl = 10) would have decoded to garbage.
    pcs_create now returns PCS_E_LUT_RANGE for such tables -- before touching CUDA, so this runs on the CPU box."""
    import ctypes as C

    import numpy as np
    from picsong_b200 import _lib
    lib = _lib.load()
    p = _lib.PcsParams()
    p.width = p.height = 2048
    p.wavelet_levels, p.lossy, p.quant_size, p.coding_passes, p.k_factor = 10, 0, 1.0, 2, 0.0
    p.cb_width, p.cb_height, p.frames = 64, 18, 0
    p.bit_depth, p.bits_per_sample, p.endianess, p.is_signed, p.components, p.is_rgb = 8, 8, 0, 0, 1, 0
    p.max_batch, p.flags = 1, 0
    nBp, nSub, cR, cG, cS = 16, 3, 1, 4, 9
    for i, v in enumerate((nBp, nSub, cR, cG, cS, 7, 3, 16)):
        p.lut_header[i] = v
    table = np.full((nSub * 10 + 1) * nBp * (cR + cS + cG), 64, dtype=np.int32)
    p.lut = table.ctypes.data_as(C.POINTER(C.c_int32))
    p.lut_len = int(table.size)
    ctx = C.c_void_p()
    assert lib.pcs_create(C.byref(p), C.byref(ctx)) == -3            # PCS_E_LUT_RANGE: (3*10+1)*16 + 32 = 528 rows > 497
    p.bit_depth = 17
    assert lib.pcs_create(C.byref(p), C.byref(ctx)) == -1            # PCS_E_ARG
