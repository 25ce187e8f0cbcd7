import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def ref_ext():
    """The reference's own CUDA extension recompiled for sm_100a (oracle/_ref), or None."""
    import torch
    if not torch.cuda.is_available():
        return None
    from oracle import build_ref_ext
    try:
        return build_ref_ext.load_module()
    except Exception as e:  # pragma: no cover - diagnostic only
        print("reference extension not loadable:", e)
        return None
