"""picsong_b200 -- Python binding (ctypes) of libpicsong_b200.so, the B200-native PICSONG hot path.

This package is plumbing for tests and benchmarks: PyTorch owns device memory, streams and
torch.distributed; every codec operation goes through the C ABI declared in
include/picsong_b200.h.  There is no CPU fallback: importing `picsong_b200.codec` without the
built library raises, and creating a context without a CUDA device fails with the CUDA error.
"""
from . import synth  # noqa: F401
