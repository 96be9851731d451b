import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cuda-image-and-video-codec_b200")
for p in (ROOT, PKG, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def lut_root(tmp_path_factory):
    """The four reference LUT folders (k=0, greyscale tables), re-materialised from the fixture."""
    import lut_fixture
    root = tmp_path_factory.mktemp("LUT")
    lut_fixture.materialize_all(str(root))
    return str(root)


@pytest.fixture(scope="session")
def golden():
    import golden_fixture
    return golden_fixture.Golden()
