import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# failure-path tests let one rank of a multi-GPU filter die: its peers give up after 3 s, not 60
os.environ.setdefault("SLAMRS_BARRIER_TIMEOUT_MS", "3000")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure); built on demand with gcc."""
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(autouse=True)
def _device_still_healthy(request):
    """A kernel fault is asynchronous and sticky: without this check it surfaces in whichever test
    touches the device next. Every gpu test ends with a synchronisation of every device."""
    yield
    if request.node.get_closest_marker("gpu") is None:
        return
    import torch
    for d in range(torch.cuda.device_count()):
        torch.cuda.synchronize(d)
