import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")
    # the C-ABI library is built in-tree and is not under version control: build it once on a clean checkout (nvcc cross-compiles
    # without a GPU); on the GPU box the snapshot already carries the .so
    lib = os.path.join(ROOT, "contextual-image-compression_b200", "libcic.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def cic():
    import cic_b200
    return cic_b200


@pytest.fixture(scope="session")
def small_cfg():
    """A small instance of the GAN codec graph (the reference is fixed at 256x256 / base 512)."""
    return {"img_shape": (64, 64, 3), "base": 32}
