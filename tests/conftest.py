import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.checker import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.checker import Ref, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref/libgcnref.so not built (needs /root/reference)")
    return Ref()
