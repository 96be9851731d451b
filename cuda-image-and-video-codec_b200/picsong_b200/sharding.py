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


# ---------------------------------------------------------------------------------------------------------------------
# Host-buffer path (no files): pinned host pixels -> ONE concatenated host codestream shared by all ranks, and back.
# This is what bench.py's multi-GPU `e2e` leg times ("pinned host pixels -> pinned host codestream incl.
# PCIe, no disk") and what tests/test_gpu_multirank.py checks byte for byte against the single-rank stream.
#
# A rank can only place its bytes in the concatenated stream once it knows the lengths of all frames before its own,
# i.e. after the other ranks have CODED theirs.  Waiting for the whole video would serialise the stream downloads
# behind the pixel uploads (PCIe is full duplex), so frames are dealt in ROUNDS ("frame f -> rank f mod G
# in batches"): round k covers the next world * c frames, rank r codes a contiguous run of c of them and keeps the
# packed result in HBM; one small all-gather of the round's lengths gives every rank its offset; the rank then issues
# the device-to-host copy STRAIGHT to its place in the shared stream (no staging copy, no host memcpy) on a copy
# stream, where it overlaps the next round's uploads and kernels.  Decoding needs no exchange at all (every rank has
# all lengths up front) and uses one contiguous chunk of frames per rank.
# ---------------------------------------------------------------------------------------------------------------------
def round_plan(total: int, world: int, frames_per_rank_round: int) -> list[list[tuple[int, int]]]:
    """rounds[k][r] = (first global frame, count) of rank r in round k."""
    c = max(1, int(frames_per_rank_round))
    rounds, start = [], 0
    while start < total:
        n = min(world * c, total - start)
        rounds.append([(start + partition(n, world, r)[0], partition(n, world, r)[1]) for r in range(world)])
        start += n
    return rounds


def local_frames(total: int, world: int, rank: int, frames_per_rank_round: int) -> list[int]:
    """Global indices of the frames rank `rank` codes, in the order its local pixel buffer holds them."""
    out = []
    for rnd in round_plan(total, world, frames_per_rank_round):
        f, n = rnd[rank]
        out.extend(range(f, f + n))
    return out


class SharedHostBuffer:
    """A byte buffer in POSIX shared memory mapped by every rank of the box and page-locked for CUDA (cudaHostRegister),
    so that each rank's copy engine writes / reads its own ranges of the ONE concatenated stream directly."""

    def __init__(self, tag: str, nbytes: int, dist=None, register=True):
        import mmap
        self.nbytes = int(nbytes)
        world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        import shutil
        import tempfile
        # RAM-backed when /dev/shm has the room (containers often cap it), else a file in the temporary directory
        base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > self.nbytes + (256 << 20) else tempfile.gettempdir()
        self.path = f"{base}/pcs_{tag}"
        self.owner = rank == 0
        if self.owner:
            with open(self.path, "wb") as f:
                f.truncate(self.nbytes)
        if world > 1:
            dist.barrier()
        self._fd = os.open(self.path, os.O_RDWR)
        self._map = mmap.mmap(self._fd, self.nbytes)
        self.array = np.frombuffer(self._map, dtype=np.uint8)
        self._registered = False
        if register:
            import torch
            rc = torch.cuda.cudart().cudaHostRegister(self.array.ctypes.data, self.nbytes, 1)  # cudaHostRegisterPortable
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister({self.path}, {self.nbytes} bytes) failed: {rc}")
            self._registered = True
        if world > 1:
            dist.barrier()

    def close(self, dist=None):
        if self._registered:
            import torch
            torch.cuda.cudart().cudaHostUnregister(self.array.ctypes.data)
            self._registered = False
        self.array = None
        try:
            self._map.close()
        except BufferError:
            pass
        os.close(self._fd)
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            dist.barrier()
        if self.owner and os.path.exists(self.path):
            os.remove(self.path)


class CudaRoundBackend:
    """Codes one round's frames with the CUDA codec (pcs_encode_frames through the C ABI, packed stream left in HBM)
    and places the result in host memory with an asynchronous copy on its own stream."""

    def __init__(self, codec, device, max_round_frames: int):
        import torch
        self.torch, self.cd, self.dev = torch, codec, device
        self.cap = max_round_frames * codec.max_words
        self.buf = [torch.empty(self.cap, dtype=torch.int16, device=device) for _ in range(2)]
        self.done = [None, None]
        self.copy_stream = torch.cuda.Stream(device=device)
        self.k = 0

    def encode(self, frames: np.ndarray, first: int):
        slot = self.k & 1
        self.k += 1
        if self.done[slot] is not None:
            self.done[slot].synchronize()  # the copy that last read this buffer
        total, lens = self.cd.encode_frames_into(frames, first, self.buf[slot].data_ptr(), self.cap)
        return slot, total, lens

    def place(self, slot: int, nwords: int, dst: np.ndarray):
        """dst: uint16 view of (page-locked) host memory, nwords long."""
        t = self.torch
        dst_t = t.from_numpy(dst.view(np.int16))
        with t.cuda.stream(self.copy_stream):
            dst_t.copy_(self.buf[slot][:nwords], non_blocking=True)
            ev = t.cuda.Event()
            ev.record(self.copy_stream)
        self.done[slot] = ev

    def wait(self):
        self.copy_stream.synchronize()

    def decode(self, words: np.ndarray, lens: np.ndarray, out: np.ndarray):
        self.cd.decode_frames(words, lens, out)


def _gather_round(lens: np.ndarray, width: int, dist, device) -> np.ndarray:
    """[world, width] int32 lengths of one round (rows padded with 0)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(lens, dtype=np.int32).reshape(1, -1)
    import torch
    world = dist.get_world_size()
    mine = torch.zeros(width, dtype=torch.int32)
    mine[: len(lens)] = torch.from_numpy(np.asarray(lens, dtype=np.int32))
    mine = mine.to(device) if device is not None else mine
    allv = torch.empty(world * width, dtype=torch.int32, device=mine.device)
    dist.all_gather_into_tensor(allv, mine)
    return allv.cpu().numpy().reshape(world, width)


def encode_rounds(backend, pixels_local: np.ndarray, total_frames: int, frames_per_rank_round: int,
                  stream_out: np.ndarray, dist=None, device=None) -> np.ndarray:
    """Every rank calls this with ITS frames (local_frames() order) in `pixels_local`; the concatenated codestream of
    all `total_frames` frames ends up in `stream_out` (uint16, shared by the ranks).  Returns the per-plane lengths of
    all frames in global order (the `_SIZE` values), identical on every rank."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    rounds = round_plan(total_frames, world, frames_per_rank_round)
    all_lens, base, lo = [], 0, 0
    for rnd in rounds:
        first, count = rnd[rank]
        width = max(n for _, n in rnd)
        slot, nwords, lens = backend.encode(pixels_local[lo:lo + count], first) if count else (0, 0, np.zeros(0, np.int32))
        lo += count
        ppf = (len(lens) // count) if count else 1  # planes per frame (3 for RGB)
        g = _gather_round(lens, width * ppf, dist, device)
        mine = base + int(g[:rank].sum())
        if int(g[rank].sum()) != nwords:
            raise RuntimeError("frame lengths do not add up to the packed stream")
        if mine + nwords > stream_out.size:
            raise RuntimeError("shared stream buffer too small")
        if count:
            backend.place(slot, nwords, stream_out[mine:mine + nwords])
        for r in range(world):
            all_lens.append(g[r, : rnd[r][1] * ppf])
        base += int(g.sum())
    backend.wait()
    if world > 1:
        dist.barrier()  # the concatenated stream is complete on the host
    return np.concatenate(all_lens).astype(np.int64)


def decode_contiguous(backend, stream: np.ndarray, all_lens: np.ndarray, pixels_out_local: np.ndarray, total_frames: int,
                      dist=None) -> tuple[int, int]:
    """Inverse: every rank decodes partition(total_frames) -- one contiguous run of frames, hence one contiguous range
    of the shared stream -- into its own pixel buffer.  Returns (first, count)."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    first, count = partition(total_frames, world, rank)
    ppf = len(all_lens) // total_frames
    off = frame_offsets(np.asarray(all_lens, dtype=np.int64))
    if count:
        a, b = int(off[first * ppf]), int(off[(first + count) * ppf])
        backend.decode(stream[a:b], np.asarray(all_lens[first * ppf:(first + count) * ppf], dtype=np.int32), pixels_out_local[:count])
    if world > 1:
        dist.barrier()
    return first, count
