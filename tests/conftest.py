import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def S():
    """The product's Python mirror (ctypes over the C-ABI library); builds the library if missing."""
    if os.environ.get("SLB200_EMUL") == "1":
        # host-logic emulation (tests/emul, CPU only): the GPU test files can be pointed at it to check the entry points' host side
        import scalapack_b200.api as api_
        api_._SO = os.path.join(ROOT, "tests", "emul", "libslb_emul.so")
    import scalapack_b200 as S_
    if not S_.have_library():
        import __graft_entry__ as g
        g.build()
    return S_


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    import oracle as O_
    O_.lib()
    return O_


@pytest.fixture(scope="session")
def ctx11(S):
    """A 1x1 BLACS grid."""
    return S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", 1, 1)
