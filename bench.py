#!/usr/bin/env python
"""bench.py -- throughput of the PICSONG encode/decode hot path on B200 (one process per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4|cfg1]

A "step" is one pass of the hot path over one batch of synthetic frames: encode (DWT -> BPC -> bit-stream
builder) followed by decode (unpack -> BPC decode -> inverse DWT) of every frame of the batch.
`value` = Mpix/s (original W*H*frames, whole job over all N GPUs) with pixels already resident in HBM;
`e2e` = the same through the host-buffer C-ABI calls (pcs_encode_frames / pcs_decode_frames) with
pinned host pixels in, host codestream out and back, host pixels out -- PCIe copies inside the timed
region.  Multi-GPU: frames are sharded across ranks (weak scaling: each rank codes its own batch),
no data-path collective; the per-frame stream lengths are exchanged with one small all-gather.

--impl reference runs the UNMODIFIED reference's own CUDA path (oracle/_ref/ref_harness: its
DWT<>::DWTEncode / BPCCuda<>::Code / ::Decode / DWTDecode stage wrappers, rebuilt for sm_100 by
oracle/build_ref.sh) on the same workload on the same GPU; the reference has no CPU implementation.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-image-and-video-codec_b200"), os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (W, H, wl, lossy, qs, lut folder, frames per step, description)
    "cfg1": (512, 512, 5, 0, 1.0, "n1_lossless", 16, "configs[0]: 512x512 8-bit grey, lossless 5/3, wl=5, 64x64 codeblocks"),
    "cfg2": (4096, 4096, 5, 1, 1.0, "n1_lossy", 32, "configs[1]: 4096x4096 8-bit grey, lossy 9/7 qs=1.0, wl=5, 64x64 codeblocks"),
    "cfg3": (7680, 4320, 6, 0, 1.0, "n1_lossless", 2, "configs[2]: 7680x4320 8-bit grey, lossless 5/3, wl=6, 64x64 codeblocks"),
    "cfg4": (3840, 2160, 5, 1, 1.0, "video_lossy", 8, "configs[3]: 3840x2160 8-bit grey video frames, lossy 9/7 qs=1, wl=5, frame-sharded"),
    # SURVEY 8f.1 (the widening row): the reference README's own example shape, 2048x2560 RGB still, lossy
    "rgb": (2048, 2560, 5, 1, 1.0, "n1_lossy", 4, "README example: 2048x2560 8-bit RGB (planar) stills, lossy 9/7 + ICT qs=1, wl=5"),
}
RGB_WORKLOADS = {"rgb"}


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


def make_frames(W, H, n, first=0, rgb=False):
    from picsong_b200 import synth
    return np.stack([(synth.synth_rgb if rgb else synth.synth_frame)(W, H, first + i) for i in range(n)])


# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from picsong_b200 import codec

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, H, wl, lossy, qs, lutname, B, desc = WORKLOADS[args.workload]
    lroot = materialize_luts()
    rgb = args.workload in RGB_WORKLOADS
    lut = codec.load_lut(f"{lroot}/{lutname}/", wl, rgb=rgb)
    os.environ["PCS_STAGE_TIMING"] = "1"
    cd = codec.Codec(codec.Params(W, H, wl, bool(lossy), qs, frames=0, max_batch=B, rgb=rgb), lut)
    frames_h = make_frames(W, H, B, first=rank * B, rgb=rgb)
    px = torch.from_numpy(frames_h).to(dev)
    stream = torch.empty(B * cd.max_words, dtype=torch.int16, device=dev)
    off = torch.empty(B * cd.planes + 1, dtype=torch.int64, device=dev)
    out = torch.empty_like(px)
    lens_all = torch.zeros(world * B * cd.planes, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > L2 (126 MB): evicts between timed iterations

    def step():
        cd.encode_device(px, rank * B, stream, off)
        if world > 1:  # the one exchange the sharded path has: per-frame stream lengths
            lens = (off[1:] - off[:-1]).to(torch.int32)
            dist.all_gather_into_tensor(lens_all, lens)
        cd.decode_device(stream, off, out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches0 = codec.kernel_launch_count()
    for _ in range(args.warmup):
        step()
    barrier()
    assert torch.equal(out, px) or lossy, "lossless round trip failed"
    launches_per_step = (codec.kernel_launch_count() - launches0) // max(args.warmup, 1)

    sampler = ClockSampler(local); sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # L2 flush, outside the events
        ev[i][0].record()
        cd.encode_device(px, rank * B, stream, off)
        if world > 1:
            dist.all_gather_into_tensor(lens_all, (off[1:] - off[:-1]).to(torch.int32))
        ev[i][1].record()
        cd.decode_device(stream, off, out)
        ev[i][2].record()
    barrier()
    t_enc = np.array([e[0].elapsed_time(e[1]) for e in ev]); t_dec = np.array([e[1].elapsed_time(e[2]) for e in ev])
    total_ms = float((t_enc + t_dec).sum())

    # ---- per-stage device times (library-side CUDA events; separate loop so the timed one has no host sync) ----
    stage = np.zeros((5, 6))
    for i in range(5):
        flush.fill_(i)
        cd.encode_device(px, rank * B, stream, off)
        stage[i, :3] = cd.last_stage_ms()
        cd.decode_device(stream, off, out)
        stage[i, 3:] = cd.last_stage_ms()

    # ---- DWT stage alone (roofline): pcs_dwt_forward over the batch, CUDA events on the launch stream ----
    coeff = torch.empty((B * cd.planes, cd.Ha, cd.Wa), dtype=torch.int32, device=dev)
    import ctypes as C
    dw = []
    for i in range(max(args.steps, 5) + 2):
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = cd.lib.pcs_dwt_forward(cd.ctx, C.c_void_p(px.data_ptr()), C.c_void_p(coeff.data_ptr()), 1, B, cd._stream())
        b.record(); torch.cuda.synchronize()
        assert rc == 0
        if i >= 2:
            dw.append(a.elapsed_time(b))
    dwt_ms = float(np.mean(dw))
    clocks = sampler.stop()

    # ---- end to end through the host-buffer API (pinned host memory, PCIe inside the timed region) ----
    pin_px = torch.from_numpy(frames_h).pin_memory()
    pin_out = torch.empty_like(pin_px).pin_memory()
    pin_stream = torch.empty(B * cd.max_words, dtype=torch.int16).pin_memory()
    hp, ho, hs = pin_px.numpy(), pin_out.numpy(), pin_stream.numpy().view(np.uint16)
    e2e_t, h2d, d2h = [], 0, 0
    for i in range(args.warmup + args.steps):
        barrier()
        t0 = time.perf_counter()
        s, lens = cd.encode_frames(hp, rank * B, hs)
        cd.decode_frames(s, lens, ho)
        if world > 1:
            dist.barrier()
        t1 = time.perf_counter()
        if i >= args.warmup:
            e2e_t.append(t1 - t0)
        h2d = hp.nbytes + s.nbytes + 8 * (B * cd.planes + 1)
        d2h = s.nbytes + ho.nbytes + 8 * (B * cd.planes + 1)
    e2e_ms = float(np.sum(e2e_t) * 1e3)

    # max over ranks
    if world > 1:
        t = torch.tensor([total_ms, e2e_ms, float(t_enc.sum()), float(t_dec.sum())], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, enc_ms, dec_ms = [float(v) for v in t.tolist()]
    else:
        enc_ms, dec_ms = float(t_enc.sum()), float(t_dec.sum())
    mpix_step = W * H * B * world / 1e6
    hbm, hbm_src = peaks()
    A = cd.A
    dwt_gbs = 5.0 * A * B * cd.planes / (dwt_ms * 1e-3) / 1e9
    res = {
        "metric": "encode+decode round-trip throughput (Mpix/s, original W*H*frames over encode time + decode time)",
        "value": mpix_step * args.steps / (total_ms * 1e-3), "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if lossy else "int32", "data": "synthetic",
        "config": {"workload": desc, "frames_per_step_per_gpu": B, "l2": "256 MiB flush write between timed iterations; working set > L2",
                   "parallelism": f"frame-sharded x{world}" if world > 1 else "single GPU",
                   **({"rank0_numa_node": numa} if numa is not None else {})},
        "encode_mpix_s": mpix_step * args.steps / (enc_ms * 1e-3), "decode_mpix_s": mpix_step * args.steps / (dec_ms * 1e-3),
        "e2e": {"value": mpix_step * args.steps / (e2e_ms * 1e-3), "unit": "Mpix/s", "h2d_bytes_per_step": int(h2d) * world,
                "d2h_bytes_per_step": int(d2h) * world},  # whole job, like `value` (every rank moves the same amount)
        "gpu_launches": int(launches_per_step * args.steps),
        "clocks": clocks,
        "roofline": {"kernel": "dwt_forward (all levels of the batch; u8 in -> 32-bit Mallat coefficients out)", "bound": "hbm",
                     "achieved": dwt_gbs, "peak": hbm, "peak_source": hbm_src, "unit": "GB/s", "frac": dwt_gbs / hbm,
                     "algorithmic_bytes_per_launch": 5 * A * B * cd.planes, "ms": dwt_ms, "traffic": None,
                     "launches_per_call": "one per wavelet level (5)"},
    }
    tr = measured_traffic()
    if tr and tr.get("frames") == B and tr.get("workload") == args.workload:
        res["roofline"]["traffic"] = tr["dram_bytes"]
        res["roofline"]["traffic_source"] = tr.get("source")
    if args.stage_detail:
        m = stage.mean(axis=0)
        res["stage_ms"] = {"enc_dwt": m[0], "enc_bpc": m[1], "enc_bsb": m[2], "dec_bsb": m[3], "dec_bpc": m[4], "dec_dwt": m[5]}
        res["bpc"] = {"bound": "integer issue rate", "encode_mpix_s": W * H * B / 1e6 / (m[1] * 1e-3), "decode_mpix_s": W * H * B / 1e6 / (m[4] * 1e-3),
                      "share_of_step": float((m[1] + m[4]) / m.sum())}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            res["cpu_baseline"] = cpu_baseline(W, H, wl, lossy, qs, f"{lroot}/{lutname}/") if not rgb else None
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


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
def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    W, H, wl, lossy, qs, lutname, B, desc = WORKLOADS[args.workload]
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    base = {"metric": "encode+decode round-trip throughput (Mpix/s, original W*H*frames over encode time + decode time)",
            "unit": "Mpix/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if lossy else "int32", "data": "synthetic", "impl": "reference",
            "config": {"workload": desc, "frames_per_step_per_gpu": B}}
    lroot = materialize_luts()
    if args.workload in RGB_WORKLOADS:
        return run_reference_rgb(args, base, lroot)
    if os.path.exists(harness):
        # the reference's own CUDA stage wrappers on this GPU, one frame per call as its API requires
        tmp = tempfile.mkdtemp(prefix="pcs_ref_")
        raw = os.path.join(tmp, "in.raw")
        make_frames(W, H, B).tofile(raw)
        js = os.path.join(tmp, "out.json")
        iters = args.warmup + args.steps  # the harness discards its first iteration as warm-up itself
        cmd = [harness, str(W), str(H), str(wl), str(lossy), repr(qs), f"{lroot}/{lutname}/", raw, str(B), str(max(iters, 2)), "-", js]
        sampler = ClockSampler(0); sampler.start()
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        clocks = sampler.stop()
        if not os.path.exists(js):
            print(json.dumps({**base, "unavailable": f"ref_harness failed rc={p.returncode}: {p.stdout[-200:]}"}))
            return
        r = json.load(open(js))
        ms = r["enc_dev_ms"] + r["dec_dev_ms"]  # per frame
        e2e_ms = r["enc_e2e_ms"] + r["dec_e2e_ms"]
        v = W * H / 1e6 / (ms * 1e-3)
        print(json.dumps({**base, "value": v, "ms_per_step": ms * B, "encode_mpix_s": r["enc_dev_mpix_s"], "decode_mpix_s": r["dec_dev_mpix_s"],
                          "stage_ms_per_frame": {"enc_dwt": r["enc_dwt_ms"], "enc_bpc_bsb": r["enc_bpc_bsb_ms"],
                                                 "dec_bpc_bsb": r["dec_bpc_bsb_ms"], "dec_dwt": r["dec_dwt_ms"]},
                          "clocks": clocks,
                          "cpu_baseline": {"value": v, "unit": "Mpix/s", "cores": 0, "kind": "reference",
                                           "sample": "the reference's own CUDA build (sm_100) on this B200; it has no CPU path"},
                          "e2e": {"value": W * H / 1e6 / (e2e_ms * 1e-3), "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    else:
        cb = cpu_baseline(W, H, wl, lossy, qs, f"{lroot}/{lutname}/")
        print(json.dumps({**base, "value": cb["value"], "cpu_baseline": cb,
                          "e2e": {"value": cb["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_reference_rgb(args, base, lroot):
    """RGB stills through the reference's own executable (its stage wrappers have no RGB entry point of their own:
    the colour transform lives in CodingEngine/DecodingEngine).  Time = what the binary itself reports as "time spent
    ... without considering allocation periods" for encode and for decode (both include its device<->host copies,
    the encode figure also its file writes to a tmpfs path); one still per run, median over the runs."""
    import re
    W, H, wl, lossy, qs, lutname, B, desc = WORKLOADS[args.workload]
    exe = os.path.join(ROOT, "oracle", "_ref", "PICSONG_ref")
    if not os.path.exists(exe):
        print(json.dumps({**base, "unavailable": "oracle/_ref/PICSONG_ref not built"}))
        return
    tmp = tempfile.mkdtemp(prefix="pcs_ref_rgb_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    raw, enc, dec = os.path.join(tmp, "in.raw"), os.path.join(tmp, "o.pcs"), os.path.join(tmp, "o.raw")
    make_frames(W, H, 1, rgb=True).tofile(raw)
    lutdir = f"{lroot}/{lutname}/"
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
                      "e2e": {"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--stage-detail", action="store_true", default=True)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
