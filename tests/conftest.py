import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu); everything else runs on CPU")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The CUDA library is compiled in-tree once per session (nvcc cross-compiles without a GPU)."""
    from laughter_detection_icsi_b200 import build
    build.build()


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback to test)")
    from laughter_detection_icsi_b200.engine import get_engine
    return get_engine(0, chunk_rows=2048)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
