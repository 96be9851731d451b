"""Multi-rank GPU test (run with `-m gpu` on a box with >= 2 B200s; skipped otherwise): the sharded product path --
CUDA codec on every rank, one NCCL all-gather of the frame lengths per round, placement of every rank's bytes in ONE
concatenated host stream / file -- must give byte-identical output to the single-rank path and to the reference
binary's own run of BASELINE configs[3]'s shape (tests/golden/ref_manifest.json: cfg4_4k_f4_wl5_q1 sha256)."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

from picsong_b200 import codec, sharding, synth  # noqa: E402


def _worker(rank, world, port, lut_root, case, tmp):
    import torch.distributed as dist
    sys.path[:0] = [p for p in os.environ["PCS_TEST_PATHS"].split(os.pathsep)]
    W, H, wl, lossy, qs, folder, total, per_round = case
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    lut = codec.load_lut(f"{lut_root}/{folder}/", wl)
    cd = codec.Codec(codec.Params(W, H, wl, bool(lossy), qs, frames=total, max_batch=4), lut)
    # ---- host-buffer path: rounds, lengths all-gathered over NCCL, bytes placed in one shared pinned stream ----
    mine = sharding.local_frames(total, world, rank, per_round)
    px = np.ascontiguousarray(np.stack([synth.synth_frame(W, H, f) for f in mine]))
    tag = f"mr{os.getppid()}_{total}_{per_round}"
    s_stream = sharding.SharedHostBuffer(tag + "_s", total * cd.max_words * 2, dist)
    s_pix = sharding.SharedHostBuffer(tag + "_p", total * W * H, dist)
    stream = s_stream.array.view(np.uint16)
    be = sharding.CudaRoundBackend(cd, dev, per_round)
    lens = sharding.encode_rounds(be, px, total, per_round, stream, dist, dev)
    first, count = sharding.partition(total, world, rank)
    out = s_pix.array.reshape(total, H, W)[first:first + count]
    sharding.decode_contiguous(be, stream, lens, out, total, dist)
    # ---- file path: contiguous chunks, pwrite of byte ranges, `_SIZE` by rank 0 (the CLI-level product path) ----
    fpath = os.path.join(tmp, f"v_{total}_{per_round}.pcs")
    flens = sharding.encode_sharded(lambda f, n: synth.synth_video(W, H, n, f), total, lambda fr, f0: cd.encode_frames(fr, f0), fpath,
                                    dist=dist, device=dev, batch=3)
    sharding.decode_sharded(fpath, flens, lambda w, l: cd.decode_frames(w, l), fpath + ".raw", W * H, dist=dist, batch=3)
    if rank == 0:
        n = int(lens.sum())
        np.savez(os.path.join(tmp, f"res_{total}_{per_round}.npz"), lens=lens, flens=flens,
                 sha_stream=hashlib.sha256(stream[:n].tobytes()).hexdigest(),
                 sha_pix=hashlib.sha256(s_pix.array.tobytes()).hexdigest(), stream_head=stream[:64].copy())
    cd.close()
    s_stream.close(dist); s_pix.close(dist)
    dist.destroy_process_group()


def _run(lut_root, tmp_path, case, world=2):
    os.environ["PCS_TEST_PATHS"] = os.pathsep.join(sys.path)
    port = 32500 + (os.getpid() + case[6] * 13 + case[7]) % 2000
    mp.spawn(_worker, args=(world, port, lut_root, case, str(tmp_path)), nprocs=world, join=True)
    return np.load(str(tmp_path / f"res_{case[6]}_{case[7]}.npz")), str(tmp_path / f"v_{case[6]}_{case[7]}.pcs")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_ranks_reproduce_the_reference_stream_of_config4(lut_root, tmp_path):
    man = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_manifest.json")))
    e = man["cfg4_4k_f4_wl5_q1"]
    res, fpath = _run(lut_root, tmp_path, (e["W"], e["H"], e["wl"], 1, e["qs"], e["lut"], 4, 1))
    assert ",".join(str(int(v)) for v in res["lens"]) == e["size_text"]
    assert str(res["sha_stream"]) == e["enc_sha256"]                        # shared-stream placement == reference file
    assert hashlib.sha256(open(fpath, "rb").read()).hexdigest() == e["enc_sha256"]   # pwrite path == reference file
    assert open(fpath + "_SIZE").read() == e["size_text"]
    # decoded pixels: same bytes from both paths, and what one rank alone decodes
    lut = codec.load_lut(f"{lut_root}/{e['lut']}/", e["wl"])
    cd = codec.Codec(codec.Params(e["W"], e["H"], e["wl"], True, e["qs"], frames=4, max_batch=4), lut)
    s1, l1 = cd.encode_frames(synth.synth_video(e["W"], e["H"], 4))
    d1 = cd.decode_frames(s1, l1)
    cd.close()
    assert hashlib.sha256(s1.tobytes()).hexdigest() == e["enc_sha256"]
    assert hashlib.sha256(d1.tobytes()).hexdigest() == str(res["sha_pix"])
    assert hashlib.sha256(open(fpath + ".raw", "rb").read()).hexdigest() == str(res["sha_pix"])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("total,per_round", [(7, 2), (5, 4)])
def test_two_ranks_equal_one_rank_ragged(lut_root, tmp_path, total, per_round):
    W, H, wl = 320, 192, 3
    res, fpath = _run(lut_root, tmp_path, (W, H, wl, 0, 1.0, "video_lossless", total, per_round))
    lut = codec.load_lut(f"{lut_root}/video_lossless/", wl)
    cd = codec.Codec(codec.Params(W, H, wl, False, 1.0, frames=total, max_batch=4), lut)
    frames = synth.synth_video(W, H, total)
    s1, l1 = cd.encode_frames(frames)
    cd.close()
    assert res["lens"].tolist() == l1.tolist() and res["flens"].tolist() == l1.tolist()
    assert str(res["sha_stream"]) == hashlib.sha256(s1.tobytes()).hexdigest()
    assert open(fpath, "rb").read() == s1.tobytes()
    assert open(fpath + "_SIZE").read() == ",".join(str(int(v)) for v in l1)
    assert str(res["sha_pix"]) == hashlib.sha256(frames.tobytes()).hexdigest()   # lossless: the input comes back
    assert open(fpath + ".raw", "rb").read() == frames.tobytes()
