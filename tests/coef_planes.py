"""Synthetic int32 coefficient planes that drive the bit-plane coder into its rare branches (test infrastructure):
the expansion fallback (a block needs more than 4095 codewords -> raw dump, size 4096; BPC/BPCEngine.cu:1904-1912,
2015-2024 of the reference) and magnitudes with bits at and above plane 16 (the 16-bit fast layout of the CUDA coder does
not hold them).  8-bit images never produce these planes, so they are fed straight into the coder stage.
Deterministic (numpy Generator, fixed seeds): the same planes go to the CPU oracle, the CUDA path and -- through
oracle/ref_harness `coef=` -- the reference's own BPCCuda::Code / ::Decode, whose outputs are committed under
tests/golden/coef_planes.npz."""
import numpy as np

W = H = 128
WL = 2

CASES = ("dense12", "dense14", "dense18", "sparse17", "sparse20", "sparse22", "mixed")


def plane(name: str) -> np.ndarray:
    rng = np.random.default_rng(abs(hash_name(name)))

    def uni(bits, shape=(H, W)):
        return rng.integers(-(1 << bits) + 1, 1 << bits, size=shape, dtype=np.int64)

    if name.startswith("dense"):
        q = uni(int(name[5:]))
    elif name.startswith("sparse"):
        q = uni(2)
        mask = rng.random((H, W)) < 0.01
        q[mask] = uni(int(name[6:]))[mask]
    elif name == "mixed":  # one block of each kind: zero, dense 13-bit (expands), sparse 2^19, ordinary 9-bit
        q = np.zeros((H, W), dtype=np.int64)
        q[:64, 64:] = uni(13, (64, 64))
        s = uni(1, (64, 64)); m = rng.random((64, 64)) < 0.02; s[m] = uni(19, (64, 64))[m]
        q[64:, :64] = s
        q[64:, 64:] = uni(9, (64, 64))
    else:
        raise KeyError(name)
    return q.astype(np.int32)


def hash_name(name: str) -> int:
    h = 2166136261
    for ch in name.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h
