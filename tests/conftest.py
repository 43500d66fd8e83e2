import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The plain-C restatement (oracle/sarw_oracle.c), built on demand."""
    import oracle_bindings as ob
    if not os.path.exists(ob.ORACLE_SO):
        ob.build_oracles()
    return ob.OracleLib()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference behind oracle/ref_harness.cpp; skipped when it was never built
    (it can only be built where /root/reference exists; the prebuilt oracle/_ref travels to the GPU box)."""
    import oracle_bindings as ob
    if not ob.RefLib.available() and os.path.isdir("/root/reference"):
        ob.build_oracles()
    if not ob.RefLib.available():
        pytest.skip("oracle/_ref not built")
    return ob.RefLib()


@pytest.fixture(scope="session")
def sarw():
    """The product binding (polymer-sarw_b200/sarw_b200.py over libsarw_b200.so). No fallback."""
    import sarw_b200
    sarw_b200.load_library()
    return sarw_b200


GOLDEN = os.path.join(ROOT, "tests", "golden")
