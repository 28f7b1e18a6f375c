import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


# every in-memory hit of the query compiler's fast cache re-prints the kernel source and compares (qs_jit.cu jit_key)
os.environ.setdefault("QSGPU_JIT_VERIFY", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import qs_oracle as O
    O.load()
    O.set_workers(min(8, os.cpu_count() or 1))
    return O


@pytest.fixture(scope="session")
def engine():
    """The device engine behind the C-ABI; fails loudly when the library or the GPU is missing."""
    from quickstep_b200 import engine as E
    E.init()
    return E


@pytest.fixture(scope="session")
def golden():
    import tpch_data as D
    return D.golden_tables()
