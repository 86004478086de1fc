import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_api

    oracle_api.build()
    oracle_api.lib()
    return oracle_api


@pytest.fixture(scope="session")
def built_lib():
    """builds libransac_b200.so if it is stale (nvcc cross-compiles without a GPU)"""
    import importlib.util

    spec = importlib.util.spec_from_file_location("rsac_build", os.path.join(ROOT, "orb-slam2-optimized_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


@pytest.fixture(scope="session")
def engine(built_lib):
    from ransac_b200 import capi

    eng = capi.Engine(0)   # raises without a CUDA device: GPU tests must not pass on a fallback
    yield eng
    eng.close()
