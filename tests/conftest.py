import os
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")
    _tune_threads()


def _tune_threads():
    """vCPU-quota'd containers run faster single-threaded; real hosts use all cores"""
    import torch
    import torch.nn.functional as F
    n = os.cpu_count() or 1
    if n == 1 or os.environ.get("OMP_NUM_THREADS"):
        return
    x, w = torch.randn(1, 64, 48, 64), torch.randn(64, 64, 3, 3)
    best, best_t = 1, None
    for nt in (1, n):
        torch.set_num_threads(nt)
        F.conv2d(x, w, padding=1)
        t = time.time()
        for _ in range(3):
            F.conv2d(x, w, padding=1)
        t = time.time() - t
        if best_t is None or t < best_t:
            best, best_t = nt, t
    torch.set_num_threads(best)
    os.environ["OMP_NUM_THREADS"] = str(best)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
