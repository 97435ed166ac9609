import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pytorch_nested-unet_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # GPU oracle settings: torch references in the tests are exact fp32 — no TF32 in cuDNN / cuBLAS, whatever order
    # the test files run in
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist before any test touches the product path."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("nunet_build", os.path.join(PKG, "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    if b.needs_build():
        b.build(verbose=False)
    yield
