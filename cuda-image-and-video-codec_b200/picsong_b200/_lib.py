"""Loads libpicsong_b200.so (built in-tree by `make -C cuda-image-and-video-codec_b200`)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PCS_LIB_PATH: developer override to A/B-test an alternative build of the same library (scripts/stage_time.py)
LIB_PATH = os.environ.get("PCS_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "libpicsong_b200.so")


class PcsParams(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("wavelet_levels", C.c_int32), ("lossy", C.c_int32),
        ("quant_size", C.c_float), ("coding_passes", C.c_int32), ("k_factor", C.c_float),
        ("cb_width", C.c_int32), ("cb_height", C.c_int32), ("frames", C.c_int32),
        ("bit_depth", C.c_int32), ("bits_per_sample", C.c_int32), ("endianess", C.c_int32),
        ("is_signed", C.c_int32), ("components", C.c_int32), ("is_rgb", C.c_int32),
        ("max_batch", C.c_int32), ("flags", C.c_uint32), ("lut_header", C.c_int32 * 8),
        ("lut", C.POINTER(C.c_int32)), ("lut_len", C.c_int32),
    ]


# every symbol include/picsong_b200.h declares: name -> (restype, argtypes)
_VP, _I32, _I64, _U16P = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_uint16)
SYMBOLS = {
    "pcs_adapted_dims": (None, [_I32, _I32, C.POINTER(_I32), C.POINTER(_I32)]),
    "pcs_adapted_dims_cb": (None, [_I32, _I32, _I32, C.POINTER(_I32), C.POINTER(_I32)]),
    "pcs_lut_load": (_I32, [C.c_char_p, _I32, C.POINTER(_I32), C.POINTER(_I32), _I32]),
    "pcs_lut_size": (_I32, [C.POINTER(_I32), _I32]),
    "pcs_lut_load_component": (_I32, [C.c_char_p, _I32, _I32, C.POINTER(_I32), C.POINTER(_I32), _I32]),
    "pcs_lut_load_k": (_I32, [C.c_char_p, _I32, _I32, C.POINTER(_I32), C.POINTER(_I32), _I32]),
    "pcs_frame_planes": (_I32, [_VP]),
    "pcs_header_pack": (None, [C.POINTER(PcsParams), _U16P]),
    "pcs_header_unpack": (_I32, [_U16P, C.POINTER(PcsParams)]),
    "pcs_frame_pixels": (C.c_size_t, [_VP]),
    "pcs_frame_codeblocks": (_I32, [_VP]),
    "pcs_frame_max_stream_words": (C.c_size_t, [_VP]),
    "pcs_create": (_I32, [C.POINTER(PcsParams), C.POINTER(_VP)]),
    "pcs_destroy": (None, [_VP]),
    "pcs_last_error": (C.c_char_p, [_VP]),
    "pcs_dwt_forward": (_I32, [_VP, _VP, _VP, _I32, _I32, _VP]),
    "pcs_dwt_inverse": (_I32, [_VP, _VP, _VP, _I32, _VP]),
    "pcs_bpc_encode": (_I32, [_VP, _VP, _I32, _VP, _VP, _I32, _VP]),
    "pcs_bpc_decode": (_I32, [_VP, _VP, _VP, _VP, _I32, _VP]),
    "pcs_bsb_pack": (_I32, [_VP, _VP, _VP, _I32, _VP, _VP, _I32, _VP]),
    "pcs_bsb_unpack": (_I32, [_VP, _VP, _VP, _VP, _VP, _I32, _VP]),
    "pcs_export_ref_codestream": (_I32, [_VP, _VP, _VP, _VP, _I32, _VP]),
    "pcs_encode_frames": (_I32, [_VP, _VP, _I32, _I32, _VP, C.c_size_t, _VP, C.POINTER(C.c_size_t)]),
    "pcs_decode_frames": (_I32, [_VP, _VP, _VP, _I32, _VP]),
    "pcs_host_alloc": (_VP, [C.c_size_t]),
    "pcs_host_free": (None, [_VP]),
    "pcs_encode_device": (_I32, [_VP, _VP, _I32, _I32, _VP, _VP, _VP]),
    "pcs_decode_device": (_I32, [_VP, _VP, _VP, _I32, _VP, _VP]),
    "pcs_last_stage_ms": (_I32, [_VP, C.POINTER(C.c_float)]),
    "pcs_kernel_launch_count": (C.c_uint64, []),
}

_lib = None


def load() -> C.CDLL:
    """Return the loaded library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C cuda-image-and-video-codec_b200` "
                "(or __graft_entry__.build()); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the ABI and the header drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
