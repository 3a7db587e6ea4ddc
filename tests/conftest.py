import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The tests exercise the real library; build it if it is not there yet."""
    lib = os.path.join(ROOT, "attentionalpoolingaction_b200", "custom_ops", "libapa_attn_pool.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
    yield
