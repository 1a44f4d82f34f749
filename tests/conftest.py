import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    """CPU-side artefacts are built on demand (seconds); the CUDA library is built by __graft_entry__.build()."""
    import shutil
    import __graft_entry__ as g
    if not os.path.exists(g.LIB) and "PB_LIB" not in os.environ and shutil.which("nvcc"):
        g.build_product()          # cross-compiles for sm_100a; the oracle's drop-in test binaries link against it
    if not os.path.exists(os.path.join(ROOT, "oracle", "libplonk_port.so")):
        g.build_oracle()
    return g


@pytest.fixture(scope="session")
def port():
    _ensure_built()
    from oracle._binding import load_port
    return load_port()


@pytest.fixture(scope="session")
def ref():
    _ensure_built()
    from oracle._binding import have_ref, load_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return load_ref()


@pytest.fixture(scope="session")
def oracle():
    """The strongest checker available: the compiled reference when oracle/_ref exists, else the C restatement."""
    _ensure_built()
    from oracle._binding import have_ref, load_port, load_ref
    return load_ref() if have_ref() else load_port()


@pytest.fixture(scope="session")
def W():
    from plonk_c_b200 import workload
    return workload


@pytest.fixture(scope="session")
def host():
    """The product's ctypes binding.  GPU tests call through it into libplonk_b200.so -- if the library is
    missing the test FAILS (no skip, no fallback)."""
    from plonk_c_b200 import host as h
    _ensure_built()
    h.lib()
    return h
