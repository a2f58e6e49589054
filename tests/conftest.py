import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def tex_pro():
    """One TextureProcessor (CUDA context + stream + plane pool) for the GPU tests."""
    import kanter_core_b200 as kc
    tp = kc.TextureProcessor.new()
    yield tp
    tp.close()


@pytest.fixture()
def tex_pro_fast():
    import kanter_core_b200 as kc
    tp = kc.TextureProcessor.new(math_mode=kc.MATH_FAST)
    yield tp
    tp.close()
