"""CPU test: the host-side list scheduler of the forward DWT's persistent work queue (csrc/dwt_fast.cu) must
emit a topological order -- a wrong order would hang the GPU, so it is checked without one."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_queue_order_is_topological(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "queue_order_check")
    csrc = os.path.join(ROOT, "cuda-image-and-video-codec_b200", "csrc")
    cmd = [nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-I", csrc, "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "queue_order_check.cu"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-1000:]
