import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LIB = os.path.join(ROOT, "attentionalpoolingaction_b200", "custom_ops", "libapa_attn_pool.so")
# test modules that never load the library: the NumPy oracle against the golden vectors and against itself
PURE_ORACLE = {"test_oracle.py", "test_oracle_backward.py"}   # (test_losses / test_eval_tail also hold GPU cases)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def _built_library():
    """Build the real library once if it is not there yet; skip (not fail) where it cannot be built."""
    if not os.path.exists(LIB):
        if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
            pytest.skip("libapa_attn_pool.so is not built and there is no nvcc to build it")
        import __graft_entry__
        __graft_entry__.build()
    yield


@pytest.fixture(autouse=True)
def _library_for_tests_that_use_it(request):
    """Every test module except the pure-oracle ones exercises the built library."""
    if os.path.basename(str(request.node.fspath)) not in PURE_ORACLE:
        request.getfixturevalue("_built_library")
    yield
