import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library (cross-compiled here, prebuilt on the GPU box)."""
    from sgufp_solver_b200 import build
    return build.build_library()


@pytest.fixture(scope="session")
def ref_available():
    from oracle import ref_dd
    if os.path.isdir("/root/reference"):
        ref_dd.build()
    return ref_dd.available()
