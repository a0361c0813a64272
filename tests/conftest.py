import os
import sys

import pytest

# The in-process multi-rank tests run one spinning kernel per rank on ONE GPU.  Streams that share a
# hardware work queue serialise (a kernel queued behind a spinning one never starts), so ask for the
# maximum number of queues BEFORE the CUDA context exists.  Irrelevant in deployment (one rank per GPU).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (tests are one of the few places allowed to use it)."""
    from oracle import oracle as O
    O.build(with_ref=os.path.isdir("/root/reference"))
    O.lib()
    return O


@pytest.fixture(scope="session")
def cos():
    """The product package over the CUDA library (built in-tree if missing)."""
    import caffeonspark_b200 as C
    if not os.path.exists(C.library_path()):
        C.build_library()
    return C


def gpu_count():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0
