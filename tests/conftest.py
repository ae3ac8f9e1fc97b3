import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200)")
    config.addinivalue_line("markers", "multigpu: test needs >= 2 CUDA devices")
    config.addinivalue_line("markers", "slow: long-running test")


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    n = _ngpu()
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    skip_multi = pytest.mark.skip(reason="needs >= 2 CUDA devices")
    for item in items:
        if "gpu" in item.keywords and n == 0:
            item.add_marker(skip_gpu)
        if "multigpu" in item.keywords and n < 2:
            item.add_marker(skip_multi)


@pytest.fixture
def local_session():
    import bigslice_b200 as bs
    sess = bs.start(bs.Local, bs.Parallelism(4))
    yield sess
    sess.shutdown()
