import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _device_count():
    try:
        from bvartools_b200._lib import lib
        return lib().bvg_device_count()
    except Exception:
        return -1


@pytest.fixture(scope="session")
def gpu():
    """The product library with a live device.  A missing library is a failure, a missing GPU a skip."""
    from bvartools_b200._lib import lib
    L = lib()                      # raises loudly if libbvargpu.so has not been built
    if L.bvg_device_count() < 1:
        pytest.skip("no CUDA device in this container (GPU tests run under gpurun)")
    return L


@pytest.fixture(scope="session")
def emu():
    import helpers
    return helpers.load_emu()
