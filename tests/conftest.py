import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_reference():
    return os.path.isdir(REFERENCE)


needs_reference = pytest.mark.skipif(not has_reference(), reason="/root/reference is not mounted here")


@pytest.fixture(scope="session")
def artifacts_dir():
    d = os.path.join(ROOT, "artifacts")
    if not os.path.exists(os.path.join(d, "t_mix.pzkp")):
        import __graft_entry__ as g
        g.build()
    return d
