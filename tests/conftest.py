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
