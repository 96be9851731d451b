"""CPU tests of the multi-GPU host logic (frame sharding + the one all-gather of stream lengths) with
world_size-2 gloo processes; the encoder is the CPU oracle so no GPU is needed."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from picsong_b200 import sharding

W, H, WL, FRAMES = 128, 128, 2, 5


def test_partition_covers_everything():
    for total in (0, 1, 5, 8, 600):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                f, c = sharding.partition(total, world, r)
                seen.extend(range(f, f + c))
            assert seen == list(range(total))
    assert sharding.band_partition(32768, 8)[3] == (12288, 4096)
    off = sharding.frame_offsets(np.array([5, 7, 9]))
    assert off.tolist() == [0, 5, 12, 21]


def _worker(rank, world, port, lut_root, out_path, dec_path):
    import torch.distributed as dist
    sys.path[:0] = [p for p in os.environ["PCS_TEST_PATHS"].split(os.pathsep)]
    from oracle import oracle as O
    from picsong_b200 import synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lut = O.lut_load(f"{lut_root}/video_lossless/", WL)

    def read(first, count):
        return synth.synth_video(W, H, count, first)

    def encode(frames, first):
        parts = [O.encode_frame(fr, WL, False, 1.0, lut, first=(first + i == 0), frames=FRAMES) for i, fr in enumerate(frames)]
        return np.concatenate(parts), np.array([p.size for p in parts], dtype=np.int32)

    lens = sharding.encode_sharded(read, FRAMES, encode, out_path, dist=dist, batch=2)

    def decode(words, ls):
        out, o = [], 0
        for n in ls:
            out.append(O.decode_frame(words[o:o + n], W, H, WL, False, 1.0, lut)); o += n
        return np.stack(out)

    sharding.decode_sharded(out_path, lens, decode, dec_path, W * H, dist=dist, batch=2)
    dist.destroy_process_group()


def test_two_rank_sharding_matches_reference_video(tmp_path, lut_root, golden):
    out_path, dec_path = str(tmp_path / "v.pcs"), str(tmp_path / "v.raw")
    os.environ["PCS_TEST_PATHS"] = os.pathsep.join(sys.path)
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, lut_root, out_path, dec_path), nprocs=2, join=True)
    from oracle import oracle as O
    from picsong_b200 import fileio, synth
    lut = O.lut_load(f"{lut_root}/video_lossless/", WL)
    want = [O.encode_frame(synth.synth_frame(W, H, f), WL, False, 1.0, lut, first=(f == 0), frames=FRAMES) for f in range(FRAMES)]
    got = np.fromfile(out_path, dtype=np.uint16)
    assert np.array_equal(got, np.concatenate(want))          # byte-identical to the single-process file
    assert fileio.read_size_sidecar(out_path).tolist() == [w.size for w in want]
    assert open(out_path + "_SIZE").read() == ",".join(str(w.size) for w in want)
    assert np.array_equal(np.fromfile(dec_path, dtype=np.uint8).reshape(FRAMES, H, W), synth.synth_video(W, H, FRAMES))
    # the first 4 frames are exactly what the reference binary produced for its 4-frame run, except that the
    # header's frame count differs (5 vs 4): compare from word 9 on
    ref = golden.stream("v128_f4_wl2_ll")
    assert np.array_equal(got[9: ref.size], ref[9:])


# ---- the host-buffer path (rounds + placement in one shared stream), world_size 2 over gloo, oracle as the codec ----
class _OracleBackend:
    def __init__(self, lut, frames_total):
        self.lut, self.frames_total, self.held = lut, frames_total, {}

    def encode(self, frames, first):
        from oracle import oracle as O
        parts = [O.encode_frame(fr, WL, False, 1.0, self.lut, first=(first + i == 0), frames=self.frames_total) for i, fr in enumerate(frames)]
        slot = len(self.held)
        self.held[slot] = np.concatenate(parts)
        return slot, self.held[slot].size, np.array([p.size for p in parts], dtype=np.int32)

    def place(self, slot, nwords, dst):
        dst[:] = self.held.pop(slot)[:nwords]

    def wait(self):
        pass

    def decode(self, words, lens, out):
        from oracle import oracle as O
        o = 0
        for i, n in enumerate(lens):
            out[i] = O.decode_frame(words[o:o + n], W, H, WL, False, 1.0, self.lut); o += n


def _round_worker(rank, world, port, lut_root, tag, total, per_round, result_path):
    import torch.distributed as dist
    sys.path[:0] = [p for p in os.environ["PCS_TEST_PATHS"].split(os.pathsep)]
    from oracle import oracle as O
    from picsong_b200 import synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lut = O.lut_load(f"{lut_root}/video_lossless/", WL)
    mine = sharding.local_frames(total, world, rank, per_round)
    px = np.stack([synth.synth_frame(W, H, f) for f in mine]) if mine else np.zeros((0, H, W), np.uint8)
    shared = sharding.SharedHostBuffer(tag, total * (W * H + 20) * 2, dist, register=False)
    stream = shared.array.view(np.uint16)
    be = _OracleBackend(lut, total)
    lens = sharding.encode_rounds(be, px, total, per_round, stream, dist)
    first, count = sharding.partition(total, world, rank)
    out = np.zeros((max(count, 1), H, W), np.uint8)
    sharding.decode_contiguous(be, stream, lens, out, total, dist)
    assert np.array_equal(out[:count], synth.synth_video(W, H, count, first))
    if rank == 0:
        np.savez(result_path, lens=lens, stream=stream[: int(lens.sum())].copy())
    shared.close(dist)
    dist.destroy_process_group()


@pytest.mark.parametrize("total,per_round", [(5, 1), (7, 2), (3, 4)])
def test_round_sharding_builds_the_single_process_stream(tmp_path, lut_root, total, per_round):
    os.environ["PCS_TEST_PATHS"] = os.pathsep.join(sys.path)
    port = 31500 + (os.getpid() + total * 7 + per_round) % 2000
    res = str(tmp_path / "r.npz")
    mp.spawn(_round_worker, args=(2, port, lut_root, f"t{os.getpid()}_{total}_{per_round}", total, per_round, res), nprocs=2, join=True)
    from oracle import oracle as O
    from picsong_b200 import synth
    lut = O.lut_load(f"{lut_root}/video_lossless/", WL)
    want = [O.encode_frame(synth.synth_frame(W, H, f), WL, False, 1.0, lut, first=(f == 0), frames=total) for f in range(total)]
    got = np.load(res)
    assert got["lens"].tolist() == [w.size for w in want]
    assert np.array_equal(got["stream"], np.concatenate(want))
