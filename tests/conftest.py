import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "multi-rtl-sdr-calibration_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    from gsmcal import build
    return build.build()


@pytest.fixture(scope="session")
def gpu(built_lib):
    import gsmcal
    if gsmcal.device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests need the B200 (there is no CPU fallback to test)")
    return gsmcal
