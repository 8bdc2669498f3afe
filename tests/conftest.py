import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def mm():
    """The product package.  The native artefacts are (re)built first when missing or older than their sources -- a fresh
    checkout has no .so (they are git-ignored); building the CPU oracle is building the checker, not using it."""
    import __graft_entry__ as g
    try:
        g.build_cuda()
        g.build_oracle()
    except Exception as e:                      # no nvcc / gcc on this host: use what is there, tests say what is missing
        print("conftest: native build skipped (%s)" % e)
    return g.load_package()
