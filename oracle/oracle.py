"""ctypes front-end of the CPU oracle (oracle/picsong_oracle.c).  TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpicsong_oracle.so")
CBSZ = 4096
_cb_rows = 64


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "picsong_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libpicsong_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def set_cb_rows(rows: int = 64):
    """Codeblock length: 64 (default) or 32 = the reference rebuilt with -DCBLOCK_LENGTH=32 (BASELINE configs[2] sweep)."""
    global CBSZ, _cb_rows
    lib().pcso_set_cb_rows(int(rows))
    _cb_rows = 32 if rows == 32 else 64
    CBSZ = 64 * _cb_rows


def set_sample_offset(offset: int = 128):
    """Level shift of the grey path: 2^(bps-1) (default 128), or 0 for -signedOrUnsigned 1."""
    lib().pcso_set_sample_offset(int(offset))


def adapted_dims(W, H):
    wa, ha = C.c_int(), C.c_int()
    lib().pcso_adapted_dims(W, H, C.byref(wa), C.byref(ha))
    return wa.value, ha.value


def pad_mirror(img: np.ndarray) -> np.ndarray:
    H, W = img.shape
    Wa, Ha = adapted_dims(W, H)
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.empty((Ha, Wa), dtype=np.uint8)
    lib().pcso_pad_mirror(_p(img, C.c_uint8), W, H, _p(out, C.c_uint8), Wa, Ha)
    return out


def dwt53_forward(px: np.ndarray, wl: int) -> np.ndarray:
    Ha, Wa = px.shape
    px = np.ascontiguousarray(px, dtype=np.uint8)
    out = np.empty((Ha, Wa), dtype=np.int32)
    assert lib().pcso_dwt53_forward(_p(px, C.c_uint8), Wa, Ha, wl, _p(out, C.c_int32)) == 0
    return out


def dwt53_inverse(q: np.ndarray, wl: int) -> np.ndarray:
    Ha, Wa = q.shape
    q = np.ascontiguousarray(q, dtype=np.int32)
    out = np.empty((Ha, Wa), dtype=np.uint8)
    assert lib().pcso_dwt53_inverse(_p(q, C.c_int32), Wa, Ha, wl, _p(out, C.c_uint8)) == 0
    return out


def dwt97_forward(px: np.ndarray, wl: int, qs: float) -> np.ndarray:
    Ha, Wa = px.shape
    px = np.ascontiguousarray(px, dtype=np.uint8)
    out = np.empty((Ha, Wa), dtype=np.float32)
    assert lib().pcso_dwt97_forward(_p(px, C.c_uint8), Wa, Ha, wl, C.c_float(qs), _p(out, C.c_float)) == 0
    return out


def dwt97_inverse(q: np.ndarray, wl: int, qs: float, want_float: bool = False):
    Ha, Wa = q.shape
    q = np.ascontiguousarray(q, dtype=np.int32)
    out = np.empty((Ha, Wa), dtype=np.uint8)
    f = np.empty((Ha, Wa), dtype=np.float32) if want_float else None
    assert lib().pcso_dwt97_inverse(_p(q, C.c_int32), Wa, Ha, wl, C.c_float(qs), _p(out, C.c_uint8),
                                    _p(f, C.c_float) if want_float else None) == 0
    return (out, f) if want_float else out


def truncate(c: np.ndarray) -> np.ndarray:
    c = np.ascontiguousarray(c, dtype=np.float32)
    q = np.empty(c.shape, dtype=np.int32)
    lib().pcso_truncate(_p(c, C.c_float), C.c_size_t(c.size), _p(q, C.c_int32))
    return q


class Lut:
    """Parsed LUT folder for one wavelet-level count: hdr (8 ints) + flat int32 table."""

    def __init__(self, hdr, table):
        self.hdr = np.ascontiguousarray(hdr, dtype=np.int32)
        self.table = np.ascontiguousarray(table, dtype=np.int32)


def lut_load(folder: str, wl: int, component: int = 0) -> Lut:
    hdr = np.zeros(8, dtype=np.int32)
    cap = 1 << 16
    table = np.zeros(cap, dtype=np.int32)
    n = lib().pcso_lut_load_component(folder.encode(), wl, component, _p(hdr, C.c_int), _p(table, C.c_int32), cap)
    if n < 0:
        raise RuntimeError(f"oracle lut_load({folder!r}, wl={wl}) failed: {n}")
    return Lut(hdr, table[:n].copy())


def lut_load_k(folder: str, wl: int, component: int = 0) -> Lut:
    """k > 0 (Engines/Engine.cu:12-100): the tables of all AMOUNT_OF_BITPLANE_FILES bit-plane files back to back."""
    hdr = np.zeros(8, dtype=np.int32)
    cap = 1 << 20
    table = np.zeros(cap, dtype=np.int32)
    n = lib().pcso_lut_load_component_k(folder.encode(), wl, component, _p(hdr, C.c_int), _p(table, C.c_int32), cap)
    if n < 0:
        raise RuntimeError(f"oracle lut_load_k({folder!r}, wl={wl}) failed: {n}")
    return Lut(hdr, table[:n].copy())


def bpc_encode(q: np.ndarray, wl: int, lut: Lut, cb0: int = 0, cb1: int = 0, k: float = 0.0):
    """k > 0: complexity-scalable mode (ENCODER only, lut from lut_load_k); exact for blocks inside one subband."""
    Ha, Wa = q.shape
    q = np.ascontiguousarray(q, dtype=np.int32)
    ncb = (Wa // 64) * (Ha // _cb_rows)
    cs = np.full(ncb * CBSZ, -1, dtype=np.int32)
    sizes = np.zeros(ncb, dtype=np.int32)
    if k > 0:
        lib().pcso_bpc_encode_k(_p(q, C.c_int32), Wa, Ha, wl, _p(lut.hdr, C.c_int), _p(lut.table, C.c_int32),
                                int(lut.table.size), C.c_float(k), _p(cs, C.c_int32), _p(sizes, C.c_int32), cb0, cb1)
        return cs, sizes
    lib().pcso_bpc_encode(_p(q, C.c_int32), Wa, Ha, wl, _p(lut.hdr, C.c_int), _p(lut.table, C.c_int32),
                          int(lut.table.size), _p(cs, C.c_int32), _p(sizes, C.c_int32), cb0, cb1)
    return cs, sizes


def bpc_decode(cs: np.ndarray, sizes: np.ndarray, Wa: int, Ha: int, wl: int, lut: Lut,
               cb0: int = 0, cb1: int = 0, k: float = 0.0) -> np.ndarray:
    cs = np.ascontiguousarray(cs, dtype=np.int32)
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    q = np.zeros((Ha, Wa), dtype=np.int32)
    if k > 0:
        lib().pcso_bpc_decode_k(_p(cs, C.c_int32), _p(sizes, C.c_int32), Wa, Ha, wl, _p(lut.hdr, C.c_int),
                                _p(lut.table, C.c_int32), int(lut.table.size), C.c_float(k), _p(q, C.c_int32), cb0, cb1)
        return q
    lib().pcso_bpc_decode(_p(cs, C.c_int32), _p(sizes, C.c_int32), Wa, Ha, wl, _p(lut.hdr, C.c_int),
                          _p(lut.table, C.c_int32), int(lut.table.size), _p(q, C.c_int32), cb0, cb1)
    return q


def header_pack(W, H, wl, lossy, qs, frames=0, components=1, cp=2, cbW=64, cbH=18, bitDepth=8,
                isRGB=0, endian=0, bps=8, isSigned=0, k=0.0) -> np.ndarray:
    out = np.zeros(9, dtype=np.uint16)
    lib().pcso_header_pack(W, H, components, cp, cbW, cbH, wl, bitDepth, int(lossy), C.c_float(qs), isRGB,
                           endian, bps, isSigned, frames, C.c_float(k), _p(out, C.c_uint16))
    return out


def header_unpack(h9: np.ndarray) -> np.ndarray:
    h9 = np.ascontiguousarray(h9, dtype=np.uint16)
    b = np.zeros(16, dtype=np.int32)
    lib().pcso_header_unpack(_p(h9, C.c_uint16), _p(b, C.c_int))
    return b


def bsb_pack(cs, sizes, hdr9=None, closed=False) -> np.ndarray:
    cs = np.ascontiguousarray(cs, dtype=np.int32)
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    n = sizes.size
    out = np.zeros(int(sizes.sum()) + n + 10, dtype=np.uint16)
    h = None if hdr9 is None else _p(np.ascontiguousarray(hdr9, dtype=np.uint16), C.c_uint16)
    fn = lib().pcso_bsb_pack_closed if closed else lib().pcso_bsb_pack
    tot = fn(_p(cs, C.c_int32), _p(sizes, C.c_int32), n, h, _p(out, C.c_uint16))
    assert tot == out.size
    return out


def bsb_unpack(stream: np.ndarray, ncb: int):
    stream = np.ascontiguousarray(stream, dtype=np.uint16)
    cs = np.empty(ncb * CBSZ, dtype=np.int32)
    sizes = np.zeros(ncb, dtype=np.int32)
    tot = lib().pcso_bsb_unpack(_p(stream, C.c_uint16), ncb, _p(cs, C.c_int32), _p(sizes, C.c_int32))
    return cs, sizes, tot


# ---- whole-frame convenience (the reference's runImage call order) -------------------------
def encode_frame(img: np.ndarray, wl: int, lossy: bool, qs: float, lut: Lut, first: bool = True,
                 frames: int = 0, k: float = 0.0) -> np.ndarray:
    """Engines/CodingEngine.cu:593-753 (grey): pad -> offset -> DWT -> BPC -> bit stream."""
    H, W = img.shape
    px = pad_mirror(img)
    if lossy:
        q = truncate(dwt97_forward(px, wl, qs))
    else:
        q = dwt53_forward(px, wl)
    cs, sizes = bpc_encode(q, wl, lut, k=k)
    hdr = header_pack(W, H, wl, lossy, qs, frames, k=k) if first else None
    return bsb_pack(cs, sizes, hdr)


def encode_frame_params(img: np.ndarray, wl: int, lossy: bool, qs: float, lut: Lut, **hdr) -> np.ndarray:
    """encode_frame with extra header fields (bitDepth, bps, isSigned ...); the level shift itself is set_sample_offset()."""
    H, W = img.shape
    px = pad_mirror(img)
    q = truncate(dwt97_forward(px, wl, qs)) if lossy else dwt53_forward(px, wl)
    cs, sizes = bpc_encode(q, wl, lut)
    return bsb_pack(cs, sizes, header_pack(W, H, wl, lossy, qs, 0, **hdr))


def decode_frame(stream: np.ndarray, Wa: int, Ha: int, wl: int, lossy: bool, qs: float, lut: Lut, k: float = 0.0) -> np.ndarray:
    """Engines/DecodingEngine.cu:765-859 (grey): unpack -> BPC decode -> inverse DWT -> clamp."""
    ncb = (Wa // 64) * (Ha // _cb_rows)
    cs, sizes, _ = bsb_unpack(stream, ncb)
    q = bpc_decode(cs, sizes, Wa, Ha, wl, lut, k=k)
    return dwt97_inverse(q, wl, qs) if lossy else dwt53_inverse(q, wl)


# ---- RGB path: colour transform, then three independently coded component planes ----
def rgb_forward(planes: np.ndarray, lossy: bool):
    """planes: u8 (3, Ha, Wa), already mirror-padded -> three transformed planes (int32 | float32): Y, B-G, R-G."""
    planes = np.ascontiguousarray(planes, dtype=np.uint8)
    n = planes[0].size
    dt, ct = (np.float32, C.c_float) if lossy else (np.int32, C.c_int32)
    out = [np.empty(planes[0].shape, dtype=dt) for _ in range(3)]
    fn = lib().pcso_ict_forward if lossy else lib().pcso_rct_forward
    fn(_p(planes[0], C.c_uint8), _p(planes[1], C.c_uint8), _p(planes[2], C.c_uint8), C.c_size_t(n),
       _p(out[0], ct), _p(out[1], ct), _p(out[2], ct))
    return out


def rgb_inverse(comps, lossy: bool) -> np.ndarray:
    dt, ct = (np.float32, C.c_float) if lossy else (np.int32, C.c_int32)
    comps = [np.ascontiguousarray(c, dtype=dt) for c in comps]
    out = np.empty((3,) + comps[0].shape, dtype=np.uint8)
    fn = lib().pcso_ict_inverse if lossy else lib().pcso_rct_inverse
    fn(_p(comps[0], ct), _p(comps[1], ct), _p(comps[2], ct), C.c_size_t(comps[0].size),
       _p(out[0], C.c_uint8), _p(out[1], C.c_uint8), _p(out[2], C.c_uint8))
    return out


def encode_rgb_frame(rgb: np.ndarray, wl: int, lossy: bool, qs: float, luts, header_on, frames: int = 0, k: float = 0.0):
    """Engines/CodingEngine.cu:598-630,680-710 (stills: component i is coded with iter = i, so only component 0
    carries the header) and :764-810 (video: iter = frame index for all three components).
    rgb: u8 (3, H, W); luts: the three component tables; header_on: 3 bools.  Returns the three streams."""
    _, H, W = rgb.shape
    padded = np.stack([pad_mirror(rgb[c]) for c in range(3)])
    comps = rgb_forward(padded, lossy)
    Ha, Wa = padded[0].shape
    streams = []
    for c in range(3):
        if lossy:
            cf = np.empty((Ha, Wa), dtype=np.float32)
            assert lib().pcso_dwt97_forward_plane(_p(comps[c], C.c_float), Wa, Ha, wl, C.c_float(qs), _p(cf, C.c_float)) == 0
            q = truncate(cf)
        else:
            q = np.empty((Ha, Wa), dtype=np.int32)
            assert lib().pcso_dwt53_forward_plane(_p(comps[c], C.c_int32), Wa, Ha, wl, _p(q, C.c_int32)) == 0
        cs, sizes = bpc_encode(q, wl, luts[c], k=k)
        hdr = header_pack(W, H, wl, lossy, qs, frames, isRGB=1, k=k) if header_on[c] else None
        streams.append(bsb_pack(cs, sizes, hdr))
    return streams


def decode_rgb_frame(streams, Wa: int, Ha: int, wl: int, lossy: bool, qs: float, luts) -> np.ndarray:
    """Engines/DecodingEngine.cu:738-770 + :599-650: three component streams -> u8 (3, Ha, Wa)."""
    ncb = (Wa // 64) * (Ha // _cb_rows)
    comps = []
    for c in range(3):
        cs, sizes, _ = bsb_unpack(streams[c], ncb)
        q = bpc_decode(cs, sizes, Wa, Ha, wl, luts[c])
        if lossy:
            f = np.empty((Ha, Wa), dtype=np.float32)
            assert lib().pcso_dwt97_inverse(_p(q, C.c_int32), Wa, Ha, wl, C.c_float(qs), None, _p(f, C.c_float)) == 0
            comps.append(f)
        else:
            o = np.empty((Ha, Wa), dtype=np.int32)
            assert lib().pcso_dwt53_inverse_plane(_p(q, C.c_int32), Wa, Ha, wl, _p(o, C.c_int32)) == 0
            comps.append(o)
    return rgb_inverse(comps, lossy)
