import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def rel_l2(a, b):
    import torch
    a = a.detach().double().flatten().cpu()
    b = b.detach().double().flatten().cpu()
    return float(torch.linalg.norm(a - b) / (torch.linalg.norm(b) + 1e-30))


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lkgd_b200 import ops
    ops.device_check(0)
    return torch.device("cuda:0")
