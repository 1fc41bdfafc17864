import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def grbda():
    """The product package; builds libgrbda_cuda.so on first use if it is missing."""
    import importlib.util
    lib = os.path.join(ROOT, "generalized_rbda_b200", "libgrbda_cuda.so")
    stamp = os.path.join(ROOT, "generalized_rbda_b200", "libgrbda_cuda.stamp")
    spec = importlib.util.spec_from_file_location("grbda_build", os.path.join(ROOT, "generalized_rbda_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    # the library under test must be built from the sources in the tree: build.py stamps it with their digest;
    # a missing or stale library is rebuilt (incremental; nvcc cross-compiles without a GPU)
    built_from = open(stamp).read().strip() if os.path.exists(stamp) else None
    if not os.path.exists(lib) or built_from != mod.source_digest():
        mod.build(verbose=False)
        assert open(stamp).read().strip() == mod.source_digest()
    import generalized_rbda_b200
    return generalized_rbda_b200


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.lib()
    return binding
