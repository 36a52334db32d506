import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
# Finished cubins are kept in the tree (git-ignored, shipped to the GPU box with the snapshot), so a test
# run does not pay NVRTC again for a scene it has compiled before (key: generated text + options + version).
os.environ.setdefault("MARAY_JIT_CACHE", os.path.join(ROOT, ".jitcache"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native pieces once per session if they are missing (nvcc cross-compiles without a GPU)."""
    so = os.path.join(ROOT, "maray_b200", "libmaray_cuda.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "maray_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


@pytest.fixture(scope="session")
def chess_bytes():
    with open(os.path.join(ROOT, "maray_b200", "data", "chess.maray"), "rb") as f:
        return f.read()
