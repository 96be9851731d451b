"""Thin object layer over the C ABI for tests / benchmarks (torch tensors in, torch tensors out).

`Codec` owns one pcs_ctx.  Stage methods take and return CUDA tensors and enqueue on the current
torch stream; `encode_frames` / `decode_frames` are the host-buffer calls a user of the reference's
CodingEngine / DecodingEngine would make (pixels in, codestream out and back).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib

PCS_F_REF_FLAT_CROP = 1
PCS_F_STILLS = 2
PCS_F_CB_LENGTH_32 = 4


class PcsError(RuntimeError):
    pass


def _check(rc: int, what: str, ctx=None):
    if rc != 0:
        msg = ""
        if ctx is not None:
            msg = (_lib.load().pcs_last_error(ctx) or b"").decode()
        kind = "CUDA error" if rc > 0 else "PCS error"
        raise PcsError(f"{what}: {kind} {rc} {msg}")


def adapted_dims(width: int, height: int, cb_rows: int = 64):
    wa, ha = C.c_int32(), C.c_int32()
    _lib.load().pcs_adapted_dims_cb(width, height, cb_rows, C.byref(wa), C.byref(ha))
    return wa.value, ha.value


@dataclass
class Lut:
    header: np.ndarray
    table: np.ndarray


def load_lut(folder: str, wl: int, rgb: bool = False, k: bool = False) -> Lut:
    """Engine::initLUT (Engines/Engine.cu:8-141 of the reference): the "R" table set for greyscale, the R, G and B sets
    back to back for the RGB path; k=True: all bit-plane files of each component (for Params(k > 0))."""
    lib = _lib.load()
    hdr = (C.c_int32 * 8)()
    cap = 1 << 20 if k else 1 << 16
    parts = []
    for comp in range(3 if rgb else 1):
        tab = (C.c_int32 * cap)()
        n = (lib.pcs_lut_load_k if k else lib.pcs_lut_load_component)(folder.encode(), wl, comp, hdr, tab, cap)
        if n <= 0:
            raise PcsError(f"pcs_lut_load_component({folder!r}, wl={wl}, {comp}) failed: {n}")
        parts.append(np.array(tab[:n], dtype=np.int32))
    return Lut(np.array(hdr[:], dtype=np.int32), np.concatenate(parts))


@dataclass
class Params:
    width: int
    height: int
    wl: int = 5
    lossy: bool = False
    qs: float = 1.0
    frames: int = 0
    max_batch: int = 1
    flags: int = 0
    cp: int = 2
    k: float = 0.0
    cb_width: int = 64
    cb_height: int = 18
    rgb: bool = False
    extra: dict = field(default_factory=dict)

    def to_c(self, lut: Lut) -> "_lib.PcsParams":
        p = _lib.PcsParams()
        p.width, p.height, p.wavelet_levels, p.lossy = self.width, self.height, self.wl, int(self.lossy)
        p.quant_size, p.coding_passes, p.k_factor = self.qs, self.cp, self.k
        p.cb_width, p.cb_height, p.frames = self.cb_width, self.cb_height, self.frames
        p.bit_depth, p.bits_per_sample, p.endianess, p.is_signed = 8, 8, 0, 0
        p.components, p.is_rgb = 1, int(self.rgb)
        for k, v in self.extra.items():
            setattr(p, k, v)
        p.max_batch, p.flags = self.max_batch, self.flags
        for i in range(8):
            p.lut_header[i] = int(lut.header[i])
        self._lut_keep = np.ascontiguousarray(lut.table, dtype=np.int32)
        p.lut = self._lut_keep.ctypes.data_as(C.POINTER(C.c_int32))
        p.lut_len = int(self._lut_keep.size)
        return p


def header_pack(params: Params, lut: Lut) -> np.ndarray:
    out = (C.c_uint16 * 9)()
    cp = params.to_c(lut)
    _lib.load().pcs_header_pack(C.byref(cp), out)
    return np.array(out[:], dtype=np.uint16)


def header_unpack(h9) -> dict:
    arr = (C.c_uint16 * 9)(*[int(v) for v in h9[:9]])
    p = _lib.PcsParams()
    rc = _lib.load().pcs_header_unpack(arr, C.byref(p))
    if rc:
        raise PcsError(f"pcs_header_unpack: {rc}")
    return {k: getattr(p, k) for k in ("width", "height", "wavelet_levels", "lossy", "quant_size", "coding_passes",
                                       "cb_width", "cb_height", "frames", "bit_depth", "bits_per_sample",
                                       "endianess", "is_signed", "components", "is_rgb", "k_factor")}


class Codec:
    def __init__(self, params: Params, lut: Lut):
        import torch  # device memory / streams only
        self.torch = torch
        self.lib = _lib.load()
        self.params = params
        self._cparams = params.to_c(lut)
        ctx = C.c_void_p()
        _check(self.lib.pcs_create(C.byref(self._cparams), C.byref(ctx)), "pcs_create")
        self.ctx = ctx
        self.A = int(self.lib.pcs_frame_pixels(ctx))
        self.ncb = int(self.lib.pcs_frame_codeblocks(ctx))
        self.max_words = int(self.lib.pcs_frame_max_stream_words(ctx))  # per frame (all its planes)
        self.planes = int(self.lib.pcs_frame_planes(ctx))               # 1 grey, 3 RGB
        self.cb_rows = 32 if (params.flags & PCS_F_CB_LENGTH_32) else 64
        self.Wa, self.Ha = adapted_dims(params.width, params.height, self.cb_rows)
        self.cbw = 64 * self.cb_rows                                    # words a codeblock owns in the sparse layout

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.pcs_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers -------------------------------------------------------------------------
    def _px_shape(self, n):
        """u8 pixels of n frames: [n, H, W] grey, [n, 3, H, W] planar RGB."""
        p = self.params
        return (n, 3, p.height, p.width) if self.planes == 3 else (n, p.height, p.width)

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    # ---- stage calls (device tensors) ----------------------------------------------------
    def dwt_forward(self, px, trunc_to_int: bool = False):
        """px: uint8 [B,H,W] cuda -> coeff [B,Ha,Wa] (int32, or float32 for lossy unless trunc_to_int)."""
        t = self.torch
        B = px.shape[0]
        dt = t.int32 if (not self.params.lossy or trunc_to_int) else t.float32
        out = t.empty((B * self.planes, self.Ha, self.Wa), dtype=dt, device=px.device)
        _check(self.lib.pcs_dwt_forward(self.ctx, self._p(px), self._p(out), int(trunc_to_int), B, self._stream()),
               "pcs_dwt_forward", self.ctx)
        return out

    def dwt_inverse(self, q):
        t = self.torch
        B = q.shape[0] // self.planes
        out = t.empty(self._px_shape(B), dtype=t.uint8, device=q.device)
        _check(self.lib.pcs_dwt_inverse(self.ctx, self._p(q), self._p(out), B, self._stream()), "pcs_dwt_inverse", self.ctx)
        return out

    def bpc_encode(self, coeff):
        t = self.torch
        B = coeff.shape[0] // self.planes
        cw = t.empty((B * self.planes, self.ncb, self.cbw), dtype=t.int16, device=coeff.device)
        sizes = t.empty((B * self.planes, self.ncb), dtype=t.int32, device=coeff.device)
        _check(self.lib.pcs_bpc_encode(self.ctx, self._p(coeff), int(coeff.dtype == t.float32), self._p(cw), self._p(sizes),
                                       B, self._stream()), "pcs_bpc_encode", self.ctx)
        return cw, sizes

    def bpc_decode(self, cw, sizes):
        t = self.torch
        B = cw.shape[0] // self.planes
        q = t.empty((B * self.planes, self.Ha, self.Wa), dtype=t.int32, device=cw.device)
        _check(self.lib.pcs_bpc_decode(self.ctx, self._p(cw), self._p(sizes), self._p(q), B, self._stream()),
               "pcs_bpc_decode", self.ctx)
        return q

    def bsb_pack(self, cw, sizes, first_frame_index: int = 0):
        t = self.torch
        B = cw.shape[0] // self.planes
        out = t.empty(B * self.max_words, dtype=t.int16, device=cw.device)
        off = t.empty(B * self.planes + 1, dtype=t.int64, device=cw.device)
        _check(self.lib.pcs_bsb_pack(self.ctx, self._p(cw), self._p(sizes), first_frame_index, self._p(out), self._p(off),
                                     B, self._stream()), "pcs_bsb_pack", self.ctx)
        return out, off

    def bsb_unpack(self, stream, off):
        t = self.torch
        B = (off.numel() - 1) // self.planes
        cw = t.empty((B * self.planes, self.ncb, self.cbw), dtype=t.int16, device=stream.device)
        sizes = t.empty((B * self.planes, self.ncb), dtype=t.int32, device=stream.device)
        _check(self.lib.pcs_bsb_unpack(self.ctx, self._p(stream), self._p(off), self._p(cw), self._p(sizes), B, self._stream()),
               "pcs_bsb_unpack", self.ctx)
        return cw, sizes

    def export_ref_codestream(self, cw, sizes):
        t = self.torch
        B = cw.shape[0] // self.planes
        ref = t.empty((B * self.planes, self.A), dtype=t.int32, device=cw.device)
        _check(self.lib.pcs_export_ref_codestream(self.ctx, self._p(cw), self._p(sizes), self._p(ref), B, self._stream()),
               "pcs_export_ref_codestream", self.ctx)
        return ref

    def encode_device(self, px, first_frame_index: int = 0, out=None, off=None):
        t = self.torch
        B = px.shape[0]
        if out is None:
            out = t.empty(B * self.max_words, dtype=t.int16, device=px.device)
        if off is None:
            off = t.empty(B * self.planes + 1, dtype=t.int64, device=px.device)
        _check(self.lib.pcs_encode_device(self.ctx, self._p(px), B, first_frame_index, self._p(out), self._p(off),
                                          self._stream()), "pcs_encode_device", self.ctx)
        return out, off

    def decode_device(self, stream, off, out=None):
        t = self.torch
        B = (off.numel() - 1) // self.planes
        if out is None:
            out = t.empty(self._px_shape(B), dtype=t.uint8, device=stream.device)
        _check(self.lib.pcs_decode_device(self.ctx, self._p(stream), self._p(off), B, self._p(out), self._stream()),
               "pcs_decode_device", self.ctx)
        return out

    def last_stage_ms(self):
        ms = (C.c_float * 3)()
        _check(self.lib.pcs_last_stage_ms(self.ctx, ms), "pcs_last_stage_ms", self.ctx)
        return [float(v) for v in ms]

    # ---- host-buffer calls ---------------------------------------------------------------
    def encode_frames(self, frames: np.ndarray, first_frame_index: int = 0, out: np.ndarray | None = None):
        """frames: uint8 [N,H,W] host array -> (stream uint16 [total], lengths int32 [N])."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        n = frames.shape[0]
        if out is None:
            out = np.empty(n * self.max_words, dtype=np.uint16)
        lens = np.zeros(n * self.planes, dtype=np.int32)  # one length per plane (the `_SIZE` sidecar values)
        total = C.c_size_t()
        _check(self.lib.pcs_encode_frames(self.ctx, frames.ctypes.data, n, first_frame_index, out.ctypes.data, out.size,
                                          lens.ctypes.data, C.byref(total)), "pcs_encode_frames", self.ctx)
        return out[: total.value], lens

    def encode_frames_into(self, frames: np.ndarray, first_frame_index: int, out_ptr: int, capacity_words: int):
        """pcs_encode_frames with a raw destination pointer (host or DEVICE memory, uint16 units) -> (total words,
        lengths).  The multi-rank path keeps the packed stream in HBM this way until the all-gather of the lengths has
        told the rank where its bytes go in the concatenated host stream (sharding.ShardedVideo)."""
        assert frames.dtype == np.uint8 and frames.flags["C_CONTIGUOUS"]
        n = frames.shape[0]
        lens = np.zeros(n * self.planes, dtype=np.int32)
        total = C.c_size_t()
        _check(self.lib.pcs_encode_frames(self.ctx, frames.ctypes.data, n, first_frame_index, C.c_void_p(out_ptr), capacity_words,
                                          lens.ctypes.data, C.byref(total)), "pcs_encode_frames", self.ctx)
        return int(total.value), lens

    def decode_frames(self, stream: np.ndarray, lens: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        stream = np.ascontiguousarray(stream, dtype=np.uint16)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        n = lens.size // self.planes
        if out is None:
            out = np.empty(self._px_shape(n), dtype=np.uint8)
        _check(self.lib.pcs_decode_frames(self.ctx, stream.ctypes.data, lens.ctypes.data, n, out.ctypes.data),
               "pcs_decode_frames", self.ctx)
        return out


def kernel_launch_count() -> int:
    return int(_lib.load().pcs_kernel_launch_count())
