"""pytest configuration: registers the `gpu` marker and puts the product package
(rt-gaussian-splat-renderer_b200/, which holds the `rtgs` Python mirror) and the repo root
(for `oracle`) on sys.path.  `-m "not gpu"` runs everywhere; `-m gpu` needs a B200."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "rt-gaussian-splat-renderer_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

DATA = ROOT / "tests" / "data"
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library, built on demand (nvcc cross-compiles without a GPU)."""
    sys.path.insert(0, str(PKG))
    import build as _build
    _build.build()
    from rtgs import _native
    return _native.load()


@pytest.fixture(scope="session")
def test_ply():
    return DATA / "test.ply"
