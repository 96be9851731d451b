"""Deterministic synthetic 8-bit greyscale inputs.

Integer-only so that C++ and numpy agree bit for bit; all hash arithmetic is uint32 with
wrap-around.  `f` is the frame index for video and 0 for stills.  A region of a larger still
can be generated in global coordinates through (x0, y0) -- config 5 generates its row bands
that way.
"""
from __future__ import annotations

import numpy as np

SEED = 20261018


def _tri(t: np.ndarray, period: int) -> np.ndarray:
    return np.abs(2 * (t % period) - period) * 255 // period


def synth_frame(width: int, height: int, f: int = 0, x0: int = 0, y0: int = 0,
                seed: int = SEED) -> np.ndarray:
    """Return a (height, width) uint8 frame of the survey's synthetic content."""
    x = (np.arange(width, dtype=np.int64) + x0)[None, :]
    y = (np.arange(height, dtype=np.int64) + y0)[:, None]
    base = (_tri(x + 3 * f, 512) + _tri(y + 2 * f, 384)) // 2
    edge = 24 * (((x >> 6) ^ (y >> 6) ^ (f >> 3)) & 1)
    xu = x.astype(np.uint32)
    yu = y.astype(np.uint32)
    h = (xu * np.uint32(73856093)) ^ (yu * np.uint32(19349663)) \
        ^ np.uint32((f * 83492791) & 0xFFFFFFFF) ^ np.uint32(seed & 0xFFFFFFFF)
    h = h ^ (h << np.uint32(13))
    h = h ^ (h >> np.uint32(17))
    h = h ^ (h << np.uint32(5))
    noise = (h & np.uint32(15)).astype(np.int64) - 8
    px = np.clip(base + edge + noise - 12, 0, 255)
    return px.astype(np.uint8)


def synth_rgb(width: int, height: int, f: int = 0) -> np.ndarray:
    """Planar (3, height, width) RGB test frame: the grey content with a per-component phase shift, so the
    three planes are correlated the way natural colour images are (exercises the colour transforms)."""
    g = synth_frame(width, height, f).astype(np.int32)
    r = synth_frame(width, height, f, x0=5, y0=3).astype(np.int32)
    b = synth_frame(width, height, f, x0=11, y0=7).astype(np.int32)
    return np.stack([np.clip((3 * g + r) // 4 + 6, 0, 255), g, np.clip((3 * g + b) // 4 - 5, 0, 255)]).astype(np.uint8)


def synth_video(width: int, height: int, frames: int, first: int = 0) -> np.ndarray:
    return np.stack([synth_frame(width, height, f) for f in range(first, first + frames)])


def synth_video_torch(width: int, height: int, frames: int, first: int = 0, device="cpu", out=None):
    """synth_video() evaluated with torch integer ops on `device` (bit-identical: the formula is integer-only).
    The 600-frame video of BASELINE configs[3] is generated on the GPU in milliseconds instead of minutes of numpy;
    `out`: optional uint8 [frames, height, width] tensor to fill (e.g. a pinned host buffer, filled frame by frame)."""
    import torch
    dev = torch.device(device)
    x = torch.arange(width, dtype=torch.int64, device=dev)[None, :]
    y = torch.arange(height, dtype=torch.int64, device=dev)[:, None]
    if out is None:
        out = torch.empty((frames, height, width), dtype=torch.uint8, device=dev)
    M = 0xFFFFFFFF

    def tri(t, period):
        return (2 * (t % period) - period).abs() * 255 // period

    hx = ((x * 73856093) & M) ^ ((y * 19349663) & M)  # uint32 wrap-around products, carried in int64
    for i in range(frames):
        f = first + i
        base = (tri(x + 3 * f, 512) + tri(y + 2 * f, 384)) // 2
        edge = 24 * (((x >> 6) ^ (y >> 6) ^ (f >> 3)) & 1)
        h = hx ^ ((f * 83492791) & M) ^ (SEED & M)
        h = h ^ ((h << 13) & M)
        h = h ^ (h >> 17)
        h = h ^ ((h << 5) & M)
        px = (base + edge + (h & 15) - 8 - 12).clamp_(0, 255).to(torch.uint8)
        out[i].copy_(px, non_blocking=True)
    return out


def zeros_frame(width: int, height: int) -> np.ndarray:
    return np.zeros((height, width), dtype=np.uint8)


def noise_frame(width: int, height: int, seed: int = SEED) -> np.ndarray:
    """Uniform random u8 (adversarial: stresses codeword counts / the expansion fallback)."""
    x = np.arange(width, dtype=np.uint32)[None, :]
    y = np.arange(height, dtype=np.uint32)[:, None]
    h = (x * np.uint32(2654435761)) ^ (y * np.uint32(40503)) ^ np.uint32(seed & 0xFFFFFFFF)
    for _ in range(2):
        h = h ^ (h << np.uint32(13))
        h = h ^ (h >> np.uint32(17))
        h = h ^ (h << np.uint32(5))
        h = h * np.uint32(1664525) + np.uint32(1013904223)
    return ((h >> np.uint32(24)) & np.uint32(255)).astype(np.uint8)
