#!/usr/bin/env python
"""bench.py -- throughput of the PICSONG encode/decode hot path on B200 (one process per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4|cfg2|cfg3|cfg1|cfg5|rgb]

Default workload = BASELINE.json configs[3]: 3840x2160 8-bit grey video, 600 frames, lossy 9/7 qs=1, wl=5 -- a FIXED
total of 600 frames whatever N is (strong scaling; the whole video fits one GPU).  A "step" is one pass of the hot path
over the whole video: encode (DWT -> BPC -> bit-stream builder) of every frame, then decode (unpack -> BPC decode ->
inverse DWT) of every frame.

  value  Mpix/s (original W*H*frames, whole job over all N GPUs) with each rank's shard of the frames already resident in
         HBM, CUDA events, max over ranks; the ranks exchange the per-frame stream lengths with one NCCL all-gather.
  e2e    the same through the host-buffer C-ABI calls (pcs_encode_frames / pcs_decode_frames): pinned host pixels in,
         ONE concatenated host codestream out (all ranks place their bytes in a shared page-locked buffer after the
         all-gather of the lengths has given them their offsets: picsong_b200/sharding.py encode_rounds), and back to
         host pixels -- PCIe copies, the all-gathers, the offset scan and the concatenation inside the timed region.
  host_copy_ceiling_gbs   what this box's host can move when all N ranks copy pinned memory at once and do nothing
         else (H2D alone, D2H alone, both directions); e2e is also reported as a fraction of that ceiling.

The other workloads keep their round-1 meaning (cfg2 = configs[1], 4096^2 lossy stills, weak scaling: every rank codes
its own batch); cfg5 = configs[4], a 32768^2 still as 8 row bands dealt to the ranks (strong scaling).
--impl reference runs the UNMODIFIED reference (oracle/_ref, rebuilt for sm_100 by oracle/build_ref.sh) on the same
workload on one B200: its stage wrappers through oracle/_ref/ref_harness and, for video, its own executable with
-numberOfStreams swept over {1,2,4,8,20}, best kept.  The reference has no CPU implementation.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-image-and-video-codec_b200"), os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: W, H, wl, lossy, qs, lut folder, frames (total if strong else per GPU), batch (frames per launch), scaling
    "cfg1": dict(W=512, H=512, wl=5, lossy=0, qs=1.0, lut="n1_lossless", frames=16, batch=16, scaling="weak",
                 desc="configs[0]: 512x512 8-bit grey, lossless 5/3, wl=5, 64x64 codeblocks"),
    "cfg2": dict(W=4096, H=4096, wl=5, lossy=1, qs=1.0, lut="n1_lossy", frames=32, batch=32, scaling="weak",
                 desc="configs[1]: 4096x4096 8-bit grey, lossy 9/7 qs=1.0, wl=5, 64x64 codeblocks"),
    "cfg3": dict(W=7680, H=4320, wl=6, lossy=0, qs=1.0, lut="n1_lossless", frames=2, batch=2, scaling="weak",
                 desc="configs[2]: 7680x4320 8-bit grey, lossless 5/3, wl=6, 64x64 codeblocks"),
    "cfg4": dict(W=3840, H=2160, wl=5, lossy=1, qs=1.0, lut="video_lossy", frames=600, batch=96, scaling="strong",
                 desc="configs[3]: 3840x2160 8-bit grey video, 600 frames, lossy 9/7 qs=1, wl=5, frame-sharded"),
    "cfg5": dict(W=32768, H=4096, wl=5, lossy=0, qs=1.0, lut="n1_lossless", frames=8, batch=2, scaling="strong", bands=True,
                 desc="configs[4]: 32768x32768 8-bit grey still as 8 row bands of 32768x4096, lossless 5/3, wl=5, band-sharded"),
    # SURVEY 8f.1 (the widening row): the reference README's own example shape, 2048x2560 RGB still, lossy
    "rgb": dict(W=2048, H=2560, wl=5, lossy=1, qs=1.0, lut="n1_lossy", frames=4, batch=4, scaling="weak", rgb=True,
                desc="README example: 2048x2560 8-bit RGB (planar) stills, lossy 9/7 + ICT qs=1, wl=5"),
}
METRIC = "encode+decode round-trip throughput (Mpix/s, original W*H*frames over encode time + decode time)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set(); self.join(timeout=5)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bind_to_gpu_numa_node(index: int):
    """Pin this rank to the CPU cores of its GPU's NUMA node BEFORE any pinned host buffer is allocated, so that
    the buffers the e2e leg copies from/to are local to the GPU's PCIe root (first-touch placement).  Best effort."""
    try:
        bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        dom, rest = bdf.split(":", 1)
        node = int(open(f"/sys/bus/pci/devices/{dom[-4:]}:{rest}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def materialize_luts():
    import lut_fixture
    root = tempfile.mkdtemp(prefix="pcs_lut_")
    lut_fixture.materialize_all(root)
    return root


def measured_traffic():
    """dram__bytes_read+write of the forward-DWT launches of one pcs_dwt_forward call, from the committed ncu
    capture of this same command (profiles/dwt_traffic.json, written by scripts/ncu_summary.py traffic)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "dwt_traffic.json")))
    except Exception:
        return None


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def gen_frames(wl_cfg, indices, device, out=None):
    """Synthetic frames with the given GLOBAL indices, generated with torch integer ops on `device`
    (bit-identical to picsong_b200.synth.synth_frame).  cfg5: index = row band of ONE 32768^2 still.  Returns u8
    [n, H, W] ([n, 3, H, W] RGB) on `device`, or fills `out` (e.g. a pinned host tensor)."""
    import torch
    from picsong_b200 import synth
    W, H = wl_cfg["W"], wl_cfg["H"]
    if wl_cfg.get("rgb"):
        t = torch.from_numpy(np.stack([synth.synth_rgb(W, H, int(i)) for i in indices])).to(device)
        if out is not None:
            out.copy_(t)
            return out
        return t
    n = len(indices)
    res = out if out is not None else torch.empty((n, H, W), dtype=torch.uint8, device=device)
    if wl_cfg.get("bands"):
        x = torch.arange(W, dtype=torch.int64, device=device)[None, :]
        for j, b in enumerate(indices):  # band b = rows b*H .. b*H+H-1 of the still (f = 0)
            y = (torch.arange(H, dtype=torch.int64, device=device) + int(b) * H)[:, None]
            M = 0xFFFFFFFF
            tri = lambda t, p: (2 * (t % p) - p).abs() * 255 // p  # noqa: E731
            base = (tri(x, 512) + tri(y, 384)) // 2
            edge = 24 * (((x >> 6) ^ (y >> 6)) & 1)
            h = ((x * 73856093) & M) ^ ((y * 19349663) & M) ^ (synth.SEED & M)
            h = h ^ ((h << 13) & M); h = h ^ (h >> 17); h = h ^ ((h << 5) & M)
            res[j].copy_((base + edge + (h & 15) - 8 - 12).clamp_(0, 255).to(torch.uint8), non_blocking=True)
        return res
    i = 0
    idx = [int(v) for v in indices]
    while i < n:  # runs of consecutive indices
        j = i
        while j + 1 < n and idx[j + 1] == idx[j] + 1:
            j += 1
        synth.synth_video_torch(W, H, j - i + 1, idx[i], device=device, out=res[i:j + 1])
        i = j + 1
    return res


def copy_ceiling(dev, dist, world, nbytes=256 << 20, reps=4):
    """Pinned-memory copy bandwidth of this box with ALL ranks copying at once and nothing else running: H2D only,
    D2H only, both directions together.  GB/s summed over the ranks (time = max over ranks)."""
    import torch
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True); h_in.fill_(1)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.ones(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def timed(fn):
        fn(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.item()) / reps

    def up():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def down():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        up(); down()

    g = nbytes * world / 1e9
    return {"h2d_only": g / timed(up), "d2h_only": g / timed(down), "both_each_direction": g / timed(both), "ranks": world,
            "bytes_per_copy": nbytes}


# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes as C

    import torch
    import torch.distributed as dist
    from picsong_b200 import codec, sharding

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(WORKLOADS[args.workload])
    if args.wl:
        cfg["wl"] = args.wl
        cfg["desc"] += f" [wl={args.wl}]"
    if args.cb_length == 32:
        cfg["desc"] += " [64x32 codeblocks]"
    W, H, wl, lossy, qs, B = cfg["W"], cfg["H"], cfg["wl"], cfg["lossy"], cfg["qs"], cfg["batch"]
    strong = cfg["scaling"] == "strong"
    rgb = bool(cfg.get("rgb"))
    total = cfg["frames"] if strong else cfg["frames"] * world      # frames of the whole job
    hdr_frames = total if (strong or args.workload == "cfg4") else 0  # value of the header's frame field
    lroot = materialize_luts()
    if args.k > 0:  # complexity-scalable mode: needs every bit-plane file of the LUT folder (oracle/_ref/LUT_all)
        cfg["desc"] += f" [k={args.k}]"
        lut = codec.load_lut(k_lut_dir(cfg["lut"]), wl, rgb=rgb, k=True)
    else:
        lut = codec.load_lut(f"{lroot}/{cfg['lut']}/", wl, rgb=rgb)
    os.environ["PCS_STAGE_TIMING"] = "1"
    try:
        cd = codec.Codec(codec.Params(W, H, wl, bool(lossy), qs, frames=hdr_frames, max_batch=B, rgb=rgb,
                                      flags=codec.PCS_F_CB_LENGTH_32 if args.cb_length == 32 else 0, k=args.k), lut)
    except codec.PcsError as e:  # e.g. 64x32 codeblocks at a depth where a level of the (32-row padded) frame has odd height
        if rank == 0:
            print(json.dumps({"metric": METRIC, "unit": "Mpix/s", "n_gpus": world, "config": {"workload": cfg["desc"]}, "unavailable": str(e)}))
        return
    first, count = sharding.partition(total, world, rank)            # this rank's frames of the device-resident leg
    if count == 0:
        raise SystemExit(f"rank {rank}: no frames to code ({total} frames over {world} ranks)")
    px = gen_frames(cfg, range(first, first + count), dev)
    batches = [(b0, min(B, count - b0)) for b0 in range(0, count, B)]
    stream = torch.empty((len(batches), B * cd.max_words), dtype=torch.int16, device=dev)
    off = torch.empty((len(batches), B * cd.planes + 1), dtype=torch.int64, device=dev)
    out = torch.empty_like(px)
    width = -(-total // world)
    lens_mine = torch.zeros(width * cd.planes, dtype=torch.int32, device=dev)
    lens_all = torch.zeros(world * width * cd.planes, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > L2 (126 MB): evicts between timed iterations

    def encode_all():
        for i, (b0, n) in enumerate(batches):
            cd.encode_device(px[b0:b0 + n], first + b0, stream[i], off[i, : n * cd.planes + 1])
            lens_mine[b0 * cd.planes:(b0 + n) * cd.planes] = (off[i, 1: n * cd.planes + 1] - off[i, : n * cd.planes]).to(torch.int32)
        if world > 1:  # the one exchange the sharded path has: per-frame stream lengths
            dist.all_gather_into_tensor(lens_all, lens_mine)

    def decode_all():
        for i, (b0, n) in enumerate(batches):
            cd.decode_device(stream[i], off[i, : n * cd.planes + 1], out[b0:b0 + n])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches0 = codec.kernel_launch_count()
    for _ in range(args.warmup):
        encode_all(); decode_all()
    barrier()
    if not lossy:
        assert torch.equal(out, px), "lossless round trip failed"
    else:
        assert (out.to(torch.int16) - px.to(torch.int16)).abs().max().item() <= 8, "lossy round trip out of range"
    launches_per_step = (codec.kernel_launch_count() - launches0) // max(args.warmup, 1)

    sampler = ClockSampler(local); sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # L2 flush, outside the events (the working set is far larger than L2 anyway)
        ev[i][0].record()
        encode_all()
        ev[i][1].record()
        decode_all()
        ev[i][2].record()
    barrier()
    t_enc = np.array([e[0].elapsed_time(e[1]) for e in ev]); t_dec = np.array([e[1].elapsed_time(e[2]) for e in ev])
    total_ms = float((t_enc + t_dec).sum())

    # ---- per-stage device times (library-side CUDA events; separate loop so the timed one has no host sync) ----
    nb0 = batches[0][1]
    stage = np.zeros((5, 6))
    for i in range(5):
        flush.fill_(i)
        cd.encode_device(px[:nb0], first, stream[0], off[0, : nb0 * cd.planes + 1])
        stage[i, :3] = cd.last_stage_ms()
        cd.decode_device(stream[0], off[0, : nb0 * cd.planes + 1], out[:nb0])
        stage[i, 3:] = cd.last_stage_ms()

    # ---- DWT stage alone (roofline): pcs_dwt_forward over one batch, CUDA events on the launch stream ----
    coeff = torch.empty((nb0 * cd.planes, cd.Ha, cd.Wa), dtype=torch.int32, device=dev)
    dw, dwi = [], []
    for i in range(max(args.steps, 5) + 2):
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = cd.lib.pcs_dwt_forward(cd.ctx, C.c_void_p(px.data_ptr()), C.c_void_p(coeff.data_ptr()), 1, nb0, cd._stream())
        b.record(); torch.cuda.synchronize()
        assert rc == 0
        flush.fill_(i & 0xFF)
        c, d = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c.record()
        rc = cd.lib.pcs_dwt_inverse(cd.ctx, C.c_void_p(coeff.data_ptr()), C.c_void_p(out.data_ptr()), nb0, cd._stream())
        d.record(); torch.cuda.synchronize()
        assert rc == 0
        if i >= 2:
            dw.append(a.elapsed_time(b)); dwi.append(c.elapsed_time(d))
    dwt_ms, idwt_ms = float(np.mean(dw)), float(np.mean(dwi))
    clocks = sampler.stop()
    del coeff

    # ---- end to end through the host-buffer API (pinned host memory, PCIe inside the timed region) ----
    words_dev = int(lens_all.sum().item()) if world > 1 else int(lens_mine.sum().item())  # whole job's stream, uint16 units
    per_round = max(1, -(-count // max(1, args.rounds))) if world > 1 else count
    mine = sharding.local_frames(total, world, rank, per_round) if world > 1 else list(range(first, first + count))
    pin_px = torch.empty((len(mine),) + tuple(px.shape[1:]), dtype=torch.uint8, pin_memory=True)
    gen_frames(cfg, mine, dev, out=pin_px) if not rgb else pin_px.copy_(gen_frames(cfg, mine, dev))
    torch.cuda.synchronize()
    pin_out = torch.empty((count,) + tuple(px.shape[1:]), dtype=torch.uint8, pin_memory=True)
    cap_words = int(words_dev * 1.02) + 4096
    if world > 1:
        shared = sharding.SharedHostBuffer(f"bench_{os.environ.get('MASTER_PORT', '0')}", cap_words * 2, dist)
        hs = shared.array.view(np.uint16)
    else:
        shared = None
        hs = torch.empty(cap_words, dtype=torch.int16, pin_memory=True).numpy().view(np.uint16)
    hp, ho = pin_px.numpy(), pin_out.numpy()
    backend = sharding.CudaRoundBackend(cd, dev, per_round) if world > 1 else None
    e2e_t, h2d, d2h = [], 0, 0
    for i in range(args.warmup + args.steps):
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            s, lens = cd.encode_frames(hp, 0, hs)
            cd.decode_frames(s, lens, ho)
            nwords = s.size
        else:
            lens = sharding.encode_rounds(backend, hp, total, per_round, hs, dist, dev)       # barrier inside
            sharding.decode_contiguous(backend, hs, lens, ho, total, dist)                     # barrier inside
            nwords = int(lens.sum())
        t1 = time.perf_counter()
        if i >= args.warmup:
            e2e_t.append(t1 - t0)
        # whole job: every pixel goes up once and comes down once, every stream word comes down once and goes up once
        h2d = total * cd.planes * W * H + nwords * 2 + 8 * (total * cd.planes + 1)
        d2h = nwords * 2 + total * cd.planes * W * H + 8 * (total * cd.planes + 1)
    if world == 1:
        assert nwords == words_dev, "host-buffer path and device path disagree on the stream length"
    e2e_ms = float(np.sum(e2e_t) * 1e3)
    if not lossy:
        assert np.array_equal(ho, gen_frames(cfg, range(first, first + count), dev).cpu().numpy()), "e2e lossless round trip failed"
    ceiling = copy_ceiling(dev, dist, world)
    if shared is not None:
        shared.close(dist)

    # max over ranks
    if world > 1:
        t = torch.tensor([total_ms, e2e_ms, float(t_enc.sum()), float(t_dec.sum()), dwt_ms, idwt_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, enc_ms, dec_ms, dwt_ms, idwt_ms = [float(v) for v in t.tolist()]
    else:
        enc_ms, dec_ms = float(t_enc.sum()), float(t_dec.sum())
    mpix_step = W * H * total / 1e6
    hbm, hbm_src = peaks()
    A = cd.A
    dwt_gbs = 5.0 * A * nb0 * cd.planes / (dwt_ms * 1e-3) / 1e9
    idwt_gbs = 5.0 * A * nb0 * cd.planes / (idwt_ms * 1e-3) / 1e9
    e2e_s_per_step = e2e_ms * 1e-3 / args.steps
    # PCIe bytes per direction per step against what the box can move in that direction with both directions busy
    ceil_frac = (max(h2d, d2h) / 1e9 / e2e_s_per_step) / ceiling["both_each_direction"]
    res = {
        "metric": METRIC, "value": mpix_step * args.steps / (total_ms * 1e-3), "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": cfg["scaling"],
        "vs_baseline": None, "dtype": "f32" if lossy else "int32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "frames_total": total, "frames_per_gpu": count, "frames_per_launch": B,
                   "l2": "working set (pixels + coefficients + streams of the shard) far larger than L2; 256 MiB flush write between timed iterations",
                   "parallelism": (f"frame-sharded x{world}: contiguous chunks device-resident; rounds of {per_round} frames per rank "
                                   f"through host buffers") if world > 1 else "single GPU",
                   **({"rank0_numa_node": numa} if numa is not None else {})},
        "encode_mpix_s": mpix_step * args.steps / (enc_ms * 1e-3), "decode_mpix_s": mpix_step * args.steps / (dec_ms * 1e-3),
        "e2e": {"value": mpix_step * args.steps / (e2e_ms * 1e-3), "unit": "Mpix/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps,
                "fraction_of_host_copy_ceiling": ceil_frac,
                "path": "pcs_encode_frames / pcs_decode_frames on pinned host buffers" +
                        ("; sharding.encode_rounds / decode_contiguous: NCCL all-gather of lengths per round, offsets, placement "
                         "in one shared page-locked stream" if world > 1 else "")},
        "host_copy_ceiling_gbs": ceiling,
        "gpu_launches": int(launches_per_step * args.steps),
        "clocks": clocks,
        "roofline": {"kernel": "dwt_forward (all levels of one batch; u8 in -> 32-bit Mallat coefficients out)", "bound": "hbm",
                     "achieved": dwt_gbs, "peak": hbm, "peak_source": hbm_src, "unit": "GB/s", "frac": dwt_gbs / hbm,
                     "algorithmic_bytes_per_launch": 5 * A * nb0 * cd.planes, "ms": dwt_ms, "traffic": None,
                     "frames_per_call": nb0,
                     "inverse": {"achieved": idwt_gbs, "frac": idwt_gbs / hbm, "ms": idwt_ms}},
    }
    tr = measured_traffic()
    if tr and tr.get("frames") == nb0 and tr.get("workload") == args.workload:
        res["roofline"]["traffic"] = tr["dram_bytes"]
        res["roofline"]["traffic_source"] = tr.get("source")
    m = stage.mean(axis=0)
    res["stage_ms"] = {"frames": nb0, "enc_dwt": m[0], "enc_bpc": m[1], "enc_bsb": m[2], "dec_bsb": m[3], "dec_bpc": m[4], "dec_dwt": m[5]}
    res["bpc"] = {"bound": "integer issue rate", "encode_mpix_s": W * H * nb0 / 1e6 / (m[1] * 1e-3), "decode_mpix_s": W * H * nb0 / 1e6 / (m[4] * 1e-3),
                  "share_of_step": float((m[1] + m[4]) / m.sum())}
    res["bpc"].update(bpc_decisions(cd, px[:nb0], m[1], m[4]))
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.k == 0 and args.cb_length == 64:
        res["cpu_baseline"] = cpu_baseline(W, H, wl, lossy, qs, f"{lroot}/{cfg['lut']}/") if not rgb else None
    cd.close()
    del px, out, stream
    torch.cuda.empty_cache()
    if world == 1 and args.workload == "cfg4" and not args.no_continuity:
        res["cfg2_continuity"] = continuity_cfg2(dev, lroot)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


def bpc_decisions(cd, px, enc_ms, dec_ms):
    """the coder's unit of work is the coded decision, D = sum_cb 4096*(MSB_cb+1) + nnz (one
    significance-or-refinement decision per coefficient per plane from its block's MSB down, one sign decision per
    non-zero coefficient).  Counted on the device from the coefficients of the batch the stage times are taken on."""
    import torch
    q = cd.dwt_forward(px, trunc_to_int=True)                              # [n, Ha, Wa] int32 indices
    n, Ha, Wa = q.shape
    r = cd.cb_rows
    blk = q.abs().reshape(n, Ha // r, r, Wa // 64, 64).amax(dim=(2, 4)).to(torch.int64)
    msb1 = torch.where(blk > 0, torch.floor(torch.log2(blk.clamp(min=1).double())).long() + 1, torch.zeros_like(blk))
    D = int((msb1 * 64 * r).sum().item() + (q != 0).sum().item())
    return {"decisions_per_batch": D, "decisions_per_pixel": D / float(n * Ha * Wa),
            "encode_mdecisions_s": D / 1e6 / (enc_ms * 1e-3), "decode_mdecisions_s": D / 1e6 / (dec_ms * 1e-3)}


def continuity_cfg2(dev, lroot, steps=5):
    """Round 1's default workload (configs[1], 32 x 4096^2 lossy stills on one GPU) so that the numbers of the two rounds
    can be read side by side: device-resident value, e2e through host buffers, forward-DWT roofline fraction."""
    import ctypes as C

    import torch
    from picsong_b200 import codec
    cfg = WORKLOADS["cfg2"]
    W, H, wl, B = cfg["W"], cfg["H"], cfg["wl"], cfg["batch"]
    cd = codec.Codec(codec.Params(W, H, wl, True, 1.0, frames=0, max_batch=B), codec.load_lut(f"{lroot}/{cfg['lut']}/", wl))
    px = gen_frames(cfg, range(B), dev)
    stream = torch.empty(B * cd.max_words, dtype=torch.int16, device=dev)
    off = torch.empty(B + 1, dtype=torch.int64, device=dev)
    out = torch.empty_like(px)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        cd.encode_device(px, 0, stream, off); cd.decode_device(stream, off, out)
    torch.cuda.synchronize()
    ts = []
    for i in range(steps):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); cd.encode_device(px, 0, stream, off); cd.decode_device(stream, off, out); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    coeff = torch.empty((B, cd.Ha, cd.Wa), dtype=torch.int32, device=dev)
    dw = []
    for i in range(steps + 2):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); cd.lib.pcs_dwt_forward(cd.ctx, C.c_void_p(px.data_ptr()), C.c_void_p(coeff.data_ptr()), 1, B, cd._stream()); b.record()
        torch.cuda.synchronize()
        if i >= 2:
            dw.append(a.elapsed_time(b))
    hp = torch.empty(px.shape, dtype=torch.uint8, pin_memory=True); hp.copy_(px)
    ho = torch.empty(px.shape, dtype=torch.uint8, pin_memory=True)
    hs = torch.empty(B * cd.max_words, dtype=torch.int16, pin_memory=True).numpy().view(np.uint16)
    te = []
    for i in range(steps + 2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s, lens = cd.encode_frames(hp.numpy(), 0, hs); cd.decode_frames(s, lens, ho.numpy())
        if i >= 2:
            te.append(time.perf_counter() - t0)
    mpix = W * H * B / 1e6
    hbm, _ = peaks()
    gbs = 5.0 * cd.A * B / (float(np.mean(dw)) * 1e-3) / 1e9
    cd.close()
    return {"workload": cfg["desc"], "frames": B, "value": mpix / (float(np.mean(ts)) * 1e-3), "e2e": mpix / float(np.mean(te)),
            "unit": "Mpix/s", "dwt_forward_ms": float(np.mean(dw)), "dwt_forward_frac": gbs / hbm}


def k_lut_dir(name):
    d = os.path.join(ROOT, "oracle", "_ref", "LUT_all", name) + "/"
    if not os.path.isdir(d):
        raise SystemExit(f"--k needs the reference's complete LUT folder (all bit-plane files) at {d}: run oracle/build_ref.sh where /root/reference exists")
    return d


def run_reference_k(args, base, cfg):
    """Complexity-scalable mode on the reference arm: its stage wrappers are only driven with k = 0 by the harness, so this
    is the reference EXECUTABLE (-k K, video mode, a bounded sample of the frames, files in /dev/shm) and its own "time
    ... without considering allocation periods" lines.  Its decoder aborts on streams whose codeblocks straddle subbands
    (tests/golden/k_factor/README.md); then only the encode side is reported."""
    W, H, wl, lossy, qs = cfg["W"], cfg["H"], cfg["wl"], cfg["lossy"], cfg["qs"]
    exe = os.path.join(ROOT, "oracle", "_ref", "PICSONG_ref")
    frames = min(cfg["frames"], 64)
    tmp = scratch_dir(3 * frames * W * H, "pcs_refk_")
    raw, enc, dec = os.path.join(tmp, "in.raw"), os.path.join(tmp, "o.pcs"), os.path.join(tmp, "o.raw")
    write_raw(cfg, frames, raw)
    lutdir = k_lut_dir(cfg["lut"])
    best = None
    for ns in (1, 8):
        oe = subprocess.run([exe, "-cd", "0", "-wl", str(wl), "-cp", "2", "-type", str(lossy), "-qs", repr(qs), "-i", raw, "-o", enc, "-cbWidth", "64",
                             "-cbHeight", "18", "-xSize", str(W), "-ySize", str(H), "-video", "1", "-frames", str(frames), "-LUTFolder", lutdir,
                             "-numberOfStreams", str(ns), "-k", repr(args.k)], capture_output=True, text=True)
        od = subprocess.run([exe, "-cd", "1", "-i", enc, "-o", dec, "-LUTFolder", lutdir, "-video", "1", "-numberOfStreams", str(ns)], capture_output=True, text=True)
        me = re.search(r"without considering allocation periods is: ([0-9.eE+-]+)", oe.stdout)
        md = re.search(r"without considering allocation periods[^:]*is: ([0-9.eE+-]+)", od.stdout)
        r = {"streams": ns, "rc": [oe.returncode, od.returncode], "enc_s": float(me.group(1)) if me else None,
             "dec_s": float(md.group(1)) if (md and od.returncode == 0) else None}
        if r["enc_s"] and (best is None or r["enc_s"] < best["enc_s"]):
            best = r
    line = {**base, "k": args.k, "sample": f"{frames} frames through the reference executable, best of -numberOfStreams 1 / 8", "cli": best}
    if best and best["enc_s"]:
        mp = W * H * frames / 1e6
        line["encode_mpix_s"] = mp / best["enc_s"]
        if best["dec_s"]:
            line["decode_mpix_s"] = mp / best["dec_s"]
            line["value"] = mp / (best["enc_s"] + best["dec_s"])
            line["e2e"] = {"value": line["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        else:
            line["unavailable"] = "the reference decoder fails on its own -k stream for this geometry; encode side only"
    else:
        line["unavailable"] = "could not run the reference executable with -k"
    print(json.dumps(line))


def cpu_baseline(W, H, wl, lossy, qs, lutdir):
    """The CPU oracle (a scalar C port; the reference has no CPU path) on a bounded sample: one frame."""
    from oracle import oracle as O
    from picsong_b200 import synth
    lut = O.lut_load(lutdir, wl)
    sw, sh = min(W, 2048), min(H, 2048)  # bounded sample: a 2048x2048 crop-sized frame of the same content
    img = synth.synth_frame(sw, sh)
    t0 = time.perf_counter()
    s = O.encode_frame(img, wl, bool(lossy), qs, lut)
    Wa, Ha = O.adapted_dims(sw, sh)
    O.decode_frame(s, Wa, Ha, wl, bool(lossy), qs, lut)
    dt = time.perf_counter() - t0
    return {"value": sw * sh / 1e6 / dt, "unit": "Mpix/s", "cores": 1, "kind": "port",
            "sample": f"one {sw}x{sh} synthetic frame, encode+decode, oracle/picsong_oracle.c (scalar C, 1 thread)"}


# --------------------------------------------------------------------------------------------------
def scratch_dir(need_bytes, prefix):
    """A directory for the reference executable's files: RAM-backed /dev/shm when it has the room (so that the reference
    is not timed on disk I/O), else the default temporary directory."""
    import shutil
    for d in ("/dev/shm", None):
        try:
            if d is None or (os.path.isdir(d) and shutil.disk_usage(d).free > need_bytes + (1 << 30)):
                return tempfile.mkdtemp(prefix=prefix, dir=d)
        except OSError:
            pass
    return tempfile.mkdtemp(prefix=prefix)


def write_raw(cfg, nframes, path):
    """nframes synthetic frames of the workload as the header-less RAW file the reference reads (IO/IOManager.ipp:72-112)."""
    import torch
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    with open(path, "wb") as f:
        for f0 in range(0, nframes, 16):
            n = min(16, nframes - f0)
            f.write(gen_frames(cfg, range(f0, f0 + n), dev).cpu().numpy().tobytes())


def reference_cli_sweep(cfg, lutdir, raw, frames, tmp, streams=(1, 2, 4, 8, 20)):
    """The reference's own executable on the whole video (file in /dev/shm), -numberOfStreams swept, its own timing
    lines parsed: "The time spent with the app without considering allocation periods is" (Engines/CodingEngine.cu:1049-1060,
    DecodingEngine twin) and "The time spent with the app is" (Launcher.cu:229-231)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "PICSONG_ref")
    if not os.path.exists(exe):
        return None
    W, H, wl = cfg["W"], cfg["H"], cfg["wl"]
    runs = []
    for ns in streams:
        enc, dec = os.path.join(tmp, f"o{ns}.pcs"), os.path.join(tmp, f"o{ns}.raw")
        t0 = time.perf_counter()
        oe = subprocess.run([exe, "-cd", "0", "-wl", str(wl), "-cp", "2", "-type", str(cfg["lossy"]), "-qs", repr(cfg["qs"]), "-i", raw, "-o", enc,
                             "-cbWidth", "64", "-cbHeight", "18", "-xSize", str(W), "-ySize", str(H), "-video", "1", "-frames", str(frames),
                             "-LUTFolder", lutdir, "-numberOfStreams", str(ns), "-k", "0"], capture_output=True, text=True)
        t1 = time.perf_counter()
        od = subprocess.run([exe, "-cd", "1", "-i", enc, "-o", dec, "-LUTFolder", lutdir, "-video", "1", "-numberOfStreams", str(ns)],
                            capture_output=True, text=True)
        t2 = time.perf_counter()
        me = re.search(r"without considering allocation periods is: ([0-9.eE+-]+)", oe.stdout)
        md = re.search(r"without considering allocation periods[^:]*is: ([0-9.eE+-]+)", od.stdout)
        ok = oe.returncode == 0 and od.returncode == 0 and me and md and os.path.exists(dec) and os.path.getsize(dec) == frames * W * H
        runs.append({"streams": ns, "ok": bool(ok), "enc_s": float(me.group(1)) if me else None, "dec_s": float(md.group(1)) if md else None,
                     "enc_wall_s": t1 - t0, "dec_wall_s": t2 - t1, "rc": [oe.returncode, od.returncode]})
        for p in (enc, enc + "_SIZE", dec):
            if os.path.exists(p):
                os.remove(p)
    good = [r for r in runs if r["ok"]]
    best = min(good, key=lambda r: r["enc_s"] + r["dec_s"]) if good else None
    return {"runs": runs, "best": best}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cfg = dict(WORKLOADS[args.workload])
    if args.wl:
        cfg["wl"] = args.wl
        cfg["desc"] += f" [wl={args.wl}]"
    if args.cb_length == 32:
        cfg["desc"] += " [64x32 codeblocks]"
    W, H, wl, lossy, qs = cfg["W"], cfg["H"], cfg["wl"], cfg["lossy"], cfg["qs"]
    video = args.workload == "cfg4"
    total = cfg["frames"]
    # 64x32 codeblocks: the reference rebuilt with -DCBLOCK_LENGTH=32 (oracle/build_ref.sh with REF_SUFFIX=_cb32)
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness" + ("_cb32" if args.cb_length == 32 else ""))
    base = {"metric": METRIC, "unit": "Mpix/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32" if lossy else "int32", "data": "synthetic", "impl": "reference",
            "config": {"workload": cfg["desc"], "frames_total": total}}
    lroot = materialize_luts()
    lutdir = f"{lroot}/{cfg['lut']}/"
    if cfg.get("rgb"):
        return run_reference_rgb(args, base, lroot)
    if args.k > 0:
        return run_reference_k(args, base, cfg)
    if not os.path.exists(harness):
        cb = cpu_baseline(W, H, wl, lossy, qs, lutdir)
        print(json.dumps({**base, "value": cb["value"], "cpu_baseline": cb,
                          "e2e": {"value": cb["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    tmp = scratch_dir((3 * total if video else 24) * W * H, "pcs_ref_")
    # (1) the reference's own CUDA stage wrappers, frame by frame as its API requires, on a bounded sample of the video
    sample = min(total, 24)
    raw = os.path.join(tmp, "in.raw")
    write_raw(cfg, total if video else sample, raw)
    js = os.path.join(tmp, "out.json")
    iters = max(2, min(args.warmup + args.steps, 4))  # the harness discards its first iteration as warm-up itself
    cmd = [harness, str(W), str(H), str(wl), str(lossy), repr(qs), lutdir, raw, str(sample if (video or cfg.get("bands")) else 0), str(iters), "-", js]
    sampler = ClockSampler(0); sampler.start()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if not os.path.exists(js):
        sampler.stop()
        print(json.dumps({**base, "unavailable": f"ref_harness failed rc={p.returncode}: {p.stdout[-200:]}"}))
        return
    r = json.load(open(js))
    if r.get("roundtrip_bad"):
        base["reference_roundtrip_bad_pixels"] = r["roundtrip_bad"]  # the reference itself does not reproduce its input here
    ms = r["enc_dev_ms"] + r["dec_dev_ms"]  # per frame
    e2e_ms = r["enc_e2e_ms"] + r["dec_e2e_ms"]
    v = W * H / 1e6 / (ms * 1e-3)
    e2e_v = W * H / 1e6 / (e2e_ms * 1e-3)
    A = -(-W // 64) * 64 * (-(-H // 64) * 64)
    mean_len = float(np.mean(r["frame_len"])) if r.get("frame_len") else 0.0
    # what the harness copies per frame: padded u8 pixels up, packed stream down; packed stream up, 32-bit pixels down
    h2d, d2h = (A + 2 * mean_len) * total, (2 * mean_len + 4 * A) * total
    line = {**base, "value": v, "ms_per_step": ms * total, "encode_mpix_s": r["enc_dev_mpix_s"], "decode_mpix_s": r["dec_dev_mpix_s"],
            "stage_ms_per_frame": {"enc_dwt": r["enc_dwt_ms"], "enc_bpc_bsb": r["enc_bpc_bsb_ms"], "dec_bpc_bsb": r["dec_bpc_bsb_ms"], "dec_dwt": r["dec_dwt_ms"]},
            "harness": {"value": v, "e2e": e2e_v, "sample": f"{sample} frames x {iters - 1} timed passes, one frame per call, default stream"}}
    # (2) video: the reference executable with its own worker-stream concurrency on the whole video
    cli = reference_cli_sweep(cfg, lutdir, raw, total, tmp) if (video and args.cb_length != 32) else None
    if cli:
        line["cli"] = cli
        if cli["best"]:
            b = cli["best"]
            cli_v = W * H * total / 1e6 / (b["enc_s"] + b["dec_s"])
            line["cli"]["value_mpix_s"] = cli_v
            line["cli"]["note"] = ("its own 'time spent ... without considering allocation periods' for encode + decode of the whole "
                                   "video, files in /dev/shm, best -numberOfStreams; includes its reader/writer threads")
            e2e_v = max(e2e_v, cli_v)
    line["clocks"] = sampler.stop()
    line["cpu_baseline"] = {"value": v, "unit": "Mpix/s", "cores": 0, "kind": "reference",
                            "sample": "the reference's own CUDA build (sm_100) on this B200; it has no CPU path"}
    line["e2e"] = {"value": e2e_v, "unit": "Mpix/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "what": "best of: stage wrappers on host buffers (harness), the executable's own pipelined run (cli)"}
    print(json.dumps(line))
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)


def run_reference_rgb(args, base, lroot):
    """RGB stills through the reference's own executable (its stage wrappers have no RGB entry point of their own:
    the colour transform lives in CodingEngine/DecodingEngine).  Time = what the binary itself reports as "time spent
    ... without considering allocation periods" for encode and for decode (both include its device<->host copies,
    the encode figure also its file writes to a tmpfs path); one still per run, median over the runs."""
    cfg = WORKLOADS[args.workload]
    W, H, wl, lossy, qs, B = cfg["W"], cfg["H"], cfg["wl"], cfg["lossy"], cfg["qs"], cfg["batch"]
    from picsong_b200 import synth
    exe = os.path.join(ROOT, "oracle", "_ref", "PICSONG_ref")
    if not os.path.exists(exe):
        print(json.dumps({**base, "unavailable": "oracle/_ref/PICSONG_ref not built"}))
        return
    tmp = tempfile.mkdtemp(prefix="pcs_ref_rgb_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    raw, enc, dec = os.path.join(tmp, "in.raw"), os.path.join(tmp, "o.pcs"), os.path.join(tmp, "o.raw")
    synth.synth_rgb(W, H, 0).tofile(raw)
    lutdir = f"{lroot}/{cfg['lut']}/"
    te, td = [], []
    sampler = ClockSampler(0); sampler.start()
    for i in range(args.warmup + args.steps):
        for p in (enc, enc + "_SIZE", dec):
            if os.path.exists(p):
                os.remove(p)
        oe = subprocess.run([exe, "-cd", "0", "-wl", str(wl), "-cp", "2", "-type", str(lossy), "-qs", repr(qs), "-i", raw, "-o", enc,
                             "-cbWidth", "64", "-cbHeight", "18", "-xSize", str(W), "-ySize", str(H), "-video", "0", "-isRGB", "1",
                             "-LUTFolder", lutdir, "-k", "0"], capture_output=True, text=True).stdout
        od = subprocess.run([exe, "-cd", "1", "-i", enc, "-o", dec, "-LUTFolder", lutdir, "-video", "0"], capture_output=True, text=True).stdout
        me = re.search(r"without considering allocation periods is: ([0-9.eE+-]+)", oe)
        md = re.search(r"without considering allocation periods and I/O is: ([0-9.eE+-]+)", od)
        if not (me and md):
            print(json.dumps({**base, "unavailable": "could not parse the reference binary's timing lines"}))
            return
        if i >= args.warmup:
            te.append(float(me.group(1))); td.append(float(md.group(1)))
    clocks = sampler.stop()
    t = float(np.median(te) + np.median(td))
    v = W * H / 1e6 / t
    print(json.dumps({**base, "value": v, "ms_per_step": t * 1e3 * B, "encode_mpix_s": W * H / 1e6 / float(np.median(te)),
                      "decode_mpix_s": W * H / 1e6 / float(np.median(td)), "clocks": clocks,
                      "cpu_baseline": {"value": v, "unit": "Mpix/s", "cores": 0, "kind": "reference",
                                       "sample": "the reference's own executable (sm_100 build) on this B200, one RGB still per run, its own reported times"},
                      "e2e": {"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 3 * W * H * B, "d2h_bytes_per_step": 3 * W * H * B}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--rounds", type=int, default=4, help="multi-GPU e2e: rounds the frames are dealt in (sharding.encode_rounds)")
    ap.add_argument("--cb-length", type=int, default=64, choices=[64, 32],
                    help="codeblock length: 64, or 32 = the reference's -DCBLOCK_LENGTH=32 build (BASELINE configs[2] sweep)")
    ap.add_argument("--wl", type=int, default=0, help="override the workload's wavelet levels")
    ap.add_argument("--k", type=float, default=0.0, help="complexity-scalable mode (-k of the reference): > 0 uses the bulk-pass coder")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-continuity", action="store_true", help="skip the extra configs[1] measurement of the default run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
