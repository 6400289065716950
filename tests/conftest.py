import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if needed) and load the CUDA library; never falls back to anything."""
    so = os.path.join(ROOT, "geoformer_b200", "libgeoformer_b200.so")
    if not os.path.exists(so):
        from geoformer_b200.build import build

        build()
    from geoformer_b200 import _capi

    return _capi.lib()
