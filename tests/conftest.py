import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/leaf_pytorch")
    skip_ref = pytest.mark.skip(reason="/root/reference not present on this machine")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


def pytest_sessionstart(session):
    """A fresh checkout has no built library (build artefacts are git-ignored): build it once so that the
    boundary tests can load the C ABI.  On the GPU box the .so travels with the snapshot and this is a no-op."""
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "leaf_pytorch_b200", "lib", "libleafk.so")
    nvcc = shutil.which("nvcc") or ("/usr/local/cuda/bin/nvcc" if os.path.isfile("/usr/local/cuda/bin/nvcc") else None)
    if not os.path.isfile(lib) and nvcc:
        subprocess.run(["sh", os.path.join(ROOT, "leaf_pytorch_b200", "csrc", "build.sh")], check=False,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
