import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fenicsx-beat_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx_factory():
    """Creates mono contexts on cuda:0 and closes them at session end.  Fails loudly without the
    built library or a GPU (there is no CPU fallback to hide behind)."""
    from beat_b200._lib import Context

    made = []

    def make(device=0):
        c = Context(device)
        made.append(c)
        return c

    yield make
    for c in made:
        c.close()
