"""CPU checks of bench.py's contract: the workloads are BASELINE.json's configs at their stated sizes, the default is the
configuration the headline metric is quoted on, and the traffic figure bench.py reports belongs to the call it times."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cuda-image-and-video-codec_b200")]


def _baseline():
    return json.load(open(os.path.join(ROOT, "BASELINE.json")))


def test_workloads_are_the_baseline_configs_at_their_stated_sizes():
    import bench
    cfgs = _baseline()["configs"]
    dims = [tuple(int(v) for v in re.search(r"(\d+)x(\d+)", c).groups()) for c in cfgs]
    w = bench.WORKLOADS
    assert (w["cfg1"]["W"], w["cfg1"]["H"]) == dims[0] and not w["cfg1"]["lossy"] and w["cfg1"]["wl"] == 5
    assert (w["cfg2"]["W"], w["cfg2"]["H"]) == dims[1] and w["cfg2"]["lossy"] and w["cfg2"]["wl"] == 5
    assert (w["cfg3"]["W"], w["cfg3"]["H"]) == dims[2] and not w["cfg3"]["lossy"] and w["cfg3"]["wl"] == 6
    assert (w["cfg4"]["W"], w["cfg4"]["H"]) == dims[3] and w["cfg4"]["lossy"] and w["cfg4"]["frames"] == 600
    assert w["cfg4"]["scaling"] == "strong"
    # configs[4]: the 32768^2 still as row bands of the full width
    assert w["cfg5"]["W"] == dims[4][0] and w["cfg5"]["H"] * w["cfg5"]["frames"] == dims[4][1] and not w["cfg5"]["lossy"]


def test_default_workload_is_the_frame_sharded_video():
    import bench
    src = open(os.path.join(ROOT, "bench.py")).read()
    m = re.search(r'add_argument\("--workload"[^)]*default="(\w+)"', src)
    assert m and m.group(1) == "cfg4", "the default bench line must be configs[3] (600 frames, strong scaling)"
    assert "600 frames" in bench.WORKLOADS["cfg4"]["desc"]


def test_committed_traffic_figure_matches_the_timed_call():
    import bench
    tr = json.load(open(os.path.join(ROOT, "profiles", "dwt_traffic.json")))
    cfg = bench.WORKLOADS[tr["workload"]]
    assert tr["workload"] == "cfg4" and tr["frames"] == cfg["batch"], "roofline.traffic must be the ncu capture of the call bench.py times"
    wa, ha = -(-cfg["W"] // 64) * 64, -(-cfg["H"] // 64) * 64
    algorithmic = 5 * wa * ha * tr["frames"]
    assert algorithmic <= tr["dram_bytes"] <= 1.25 * algorithmic, "DRAM traffic far above the algorithmic bytes means wasted re-reads"
