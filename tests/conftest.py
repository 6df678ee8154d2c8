import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_library():
    """libdynoed_cuda.so, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    from dynamicoed.jl_b200 import build, runtime
    import shutil
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        build.build_default(regenerate=False)      # no-op when the .so is newer than its sources
    return runtime.LIB_PATH
