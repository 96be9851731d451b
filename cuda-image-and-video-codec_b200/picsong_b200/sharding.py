"""Frame / row-band sharding across the GPUs of one box.

The codec is intra-only: a frame's (or band's) codestream depends on nothing outside it except
"am I global frame 0" (header present vs 0xFFFF x 9).  So each rank codes a contiguous chunk of frames
with no data-path collective; the ONLY exchange is one small all-gather of the per-frame stream
lengths (int32, units of uint16), from which every rank derives the byte offset of each of its frames
in the concatenated output.  Ranks then write their own byte ranges (pwrite) and rank 0 writes the
`_SIZE` sidecar.  The backend is torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests);
the encoder is passed in, so the host logic is testable without a GPU.
"""
from __future__ import annotations

import os
from typing import Callable, Sequence

import numpy as np


def partition(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous chunk [first, first+count) of `total` frames for `rank` (earlier ranks get the remainder)."""
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def band_partition(height: int, bands: int) -> list[tuple[int, int]]:
    """Row bands of a huge still (config 5: 32768 rows -> 8 bands of 4096): (first_row, rows) per band.
    A band is coded as an independent 'frame', exactly like the reference's video mode over the same
    row-major file (SURVEY note 6)."""
    if height % bands:
        raise ValueError("band height must divide the image height")
    h = height // bands
    return [(b * h, h) for b in range(bands)]


def gather_lengths(local_lengths: Sequence[int], total: int, dist=None, device=None) -> np.ndarray:
    """All ranks' per-frame lengths in global frame order.  One all-gather of `max chunk` int32 per rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local_lengths, dtype=np.int64)
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    width = (total + world - 1) // world
    mine = torch.zeros(width, dtype=torch.int32, device=device)
    if len(local_lengths):
        mine[: len(local_lengths)] = torch.as_tensor(np.asarray(local_lengths, dtype=np.int32), device=device)
    allv = torch.empty(world * width, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(allv, mine)
    allv = allv.cpu().numpy().reshape(world, width)
    out = []
    for r in range(world):
        _, cnt = partition(total, world, r)
        out.append(allv[r, :cnt])
    return np.concatenate(out).astype(np.int64)


def frame_offsets(lengths: np.ndarray) -> np.ndarray:
    """Exclusive scan: offset (uint16 units) of every frame in the concatenated stream, plus the total."""
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(lengths, out=off[1:])
    return off


def encode_sharded(read_frames: Callable[[int, int], np.ndarray], total_frames: int,
                   encode: Callable[[np.ndarray, int], tuple[np.ndarray, np.ndarray]], out_path: str,
                   dist=None, device=None, batch: int = 16) -> np.ndarray:
    """Encode `total_frames` frames sharded over the ranks of `dist` into one file + `_SIZE` sidecar.

    read_frames(first, count) -> uint8 [count,H,W];  encode(frames, first_global_index) ->
    (uint16 stream, int32 lengths).  Returns the global length array (same on every rank)."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    first, count = partition(total_frames, world, rank)
    streams, lens = [], []
    for f0 in range(first, first + count, batch):
        n = min(batch, first + count - f0)
        s, l = encode(read_frames(f0, n), f0)
        streams.append(np.asarray(s, dtype=np.uint16)); lens.extend(int(v) for v in l)
    all_lens = gather_lengths(lens, total_frames, dist, device)
    off = frame_offsets(all_lens)
    if rank == 0:
        for p in (out_path, out_path + "_SIZE"):
            if os.path.exists(p):
                os.remove(p)
        with open(out_path, "wb") as f:
            f.truncate(int(off[-1]) * 2)
    if world > 1:
        dist.barrier()
    fd = os.open(out_path, os.O_WRONLY)
    try:
        if streams:
            os.pwrite(fd, np.concatenate(streams).tobytes(), int(off[first]) * 2)
    finally:
        os.close(fd)
    if world > 1:
        dist.barrier()
    if rank == 0:
        with open(out_path + "_SIZE", "w") as f:
            f.write(",".join(str(int(v)) for v in all_lens))
    return all_lens


def decode_sharded(stream_path: str, lengths: np.ndarray, decode: Callable[[np.ndarray, np.ndarray], np.ndarray],
                   out_path: str, frame_bytes: int, dist=None, batch: int = 16) -> None:
    """Inverse: lengths come from the `_SIZE` sidecar (IO/IOManager.ipp:196-208); every rank decodes its
    chunk of frames and writes its own byte range of the raw output."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    total = len(lengths)
    off = frame_offsets(np.asarray(lengths, dtype=np.int64))
    first, count = partition(total, world, rank)
    if rank == 0:
        if os.path.exists(out_path):
            os.remove(out_path)
        with open(out_path, "wb") as f:
            f.truncate(total * frame_bytes)
    if world > 1:
        dist.barrier()
    fd = os.open(out_path, os.O_WRONLY)
    try:
        with open(stream_path, "rb") as f:
            for f0 in range(first, first + count, batch):
                n = min(batch, first + count - f0)
                f.seek(int(off[f0]) * 2)
                words = np.frombuffer(f.read(int(off[f0 + n] - off[f0]) * 2), dtype=np.uint16)
                px = decode(words, np.asarray(lengths[f0:f0 + n], dtype=np.int32))
                os.pwrite(fd, np.ascontiguousarray(px, dtype=np.uint8).tobytes(), f0 * frame_bytes)
    finally:
        os.close(fd)
    if world > 1:
        dist.barrier()
