import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Native libraries are built in-tree by __graft_entry__.build(); build on demand if absent."""
    from vexcl_b200 import build
    build.build_lib()
    build.build_oracle()
    return True


@pytest.fixture(scope="session")
def ctx1(built):
    import vexcl_b200 as vx
    return vx.Context([0])


@pytest.fixture(scope="session")
def ctx2(built):
    """Two partition slots on the same device: the reference's own trick for exercising the
    multi-device paths on a single-GPU machine (tests/context_setup.hpp:24-39)."""
    import vexcl_b200 as vx
    return vx.Context([0, 0])


@pytest.fixture(scope="session")
def ctx3(built):
    import vexcl_b200 as vx
    return vx.Context([0, 0, 0])


@pytest.fixture(params=["ctx1", "ctx2", "ctx3"])
def ctx(request):
    return request.getfixturevalue(request.param)
