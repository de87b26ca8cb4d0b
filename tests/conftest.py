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


def _cuda_device_count() -> int:
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a host without a CUDA device, so a plain `pytest tests` works
    everywhere; on the B200 box nothing is skipped."""
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (the render path has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
