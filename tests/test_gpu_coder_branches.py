"""GPU parity of the coder's rare branches: expansion fallback (size == 4096) and bit planes >= 16,
driven by synthetic coefficient planes fed straight into pcs_bpc_encode / pcs_bpc_decode through the C ABI, against
the CPU oracle and against the reference's own BPCCuda::Code / ::Decode outputs (tests/golden/coef_planes.npz)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import coef_planes  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker only)
from picsong_b200 import codec  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "coef_planes.npz")


@pytest.fixture(scope="module")
def setup(lut_root):
    olut = O.lut_load(f"{lut_root}/n1_lossless/", coef_planes.WL)
    plut = codec.load_lut(f"{lut_root}/n1_lossless/", coef_planes.WL)
    cd = codec.Codec(codec.Params(coef_planes.W, coef_planes.H, coef_planes.WL, False, 1.0), plut)
    yield olut, cd
    cd.close()


@pytest.mark.parametrize("name", coef_planes.CASES)
def test_coder_branches_against_oracle_and_reference(name, setup):
    olut, cd = setup
    W, H, wl = coef_planes.W, coef_planes.H, coef_planes.WL
    q = coef_planes.plane(name)
    ocs, osz = O.bpc_encode(q, wl, olut)
    oq = O.bpc_decode(ocs, osz, W, H, wl, olut)
    qd = torch.from_numpy(q).cuda()[None]
    cw, sizes = cd.bpc_encode(qd)
    gsz = sizes[0].cpu().numpy()
    assert np.array_equal(gsz, osz), (name, gsz, osz)
    gcs = cd.export_ref_codestream(cw, sizes)[0].cpu().numpy()
    assert np.array_equal(gcs, ocs), f"{name}: sparse codestream differs from the oracle"
    # packed stream round trip and decode
    stream, off = cd.bsb_pack(cw, sizes, 0)
    cw_b, sizes_b = cd.bsb_unpack(stream, off)
    assert torch.equal(sizes_b, sizes)
    gq = cd.bpc_decode(cw_b, sizes_b)[0].cpu().numpy()
    assert np.array_equal(gq, oq), f"{name}: decoded coefficients differ from the oracle"
    expanded = gsz == 4096
    msb = np.array([int(np.abs(q[64 * (b // 2):64 * (b // 2) + 64, 64 * (b % 2):64 * (b % 2) + 64].astype(np.int64)).max()).bit_length() - 1
                    for b in range(4)])
    # what the format guarantees: exact unless a block was dumped raw with magnitudes >= 2^15 (the dump keeps 16 bits
    # of |q| << 1 | sign, BPCEngine.cu:1904-1912)
    for b in range(4):
        blk = (slice(64 * (b // 2), 64 * (b // 2) + 64), slice(64 * (b % 2), 64 * (b % 2) + 64))
        if not expanded[b] or msb[b] < 15:
            assert np.array_equal(gq[blk], q[blk]), (name, b)
    # the branches this test exists for must really occur
    if name in ("dense14", "dense18", "mixed"):
        assert expanded.any()
    if name in ("sparse17", "sparse20", "sparse22", "mixed"):
        assert ((msb >= 16) & ~expanded).any()
    # the reference's own coder / decoder on the same plane (oracle/ref_harness coef=..., run on a B200)
    if os.path.exists(GOLD):
        g = np.load(GOLD)
        if f"{name}.stream" in g:
            pstream = stream[: int(off.cpu().numpy()[1])].cpu().numpy().view(np.uint16)
            assert np.array_equal(pstream, g[f"{name}.stream"]), f"{name}: packed stream differs from the reference's"
            assert np.array_equal(gq, g[f"{name}.q"]), f"{name}: decoded plane differs from the reference decoder's"


def test_branch_planes_cover_both_paths():
    """host-only sanity of the fixture itself (runs under -m gpu with the rest of this file)"""
    assert np.abs(coef_planes.plane("sparse20")).max() >= 1 << 16
    assert np.abs(coef_planes.plane("sparse22")).max() < 1 << 23
