import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """No test may hang a run: 15 minutes per test (pytest-timeout, when it is installed) unless the test sets its own limit.  A
    multi-process test whose peer died would otherwise wait in a collective for ever."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for it in items:
        if it.get_closest_marker("timeout") is None:
            it.add_marker(pytest.mark.timeout(900))


@pytest.fixture(scope="session")
def built_lib():
    """Builds (if needed) and loads libb200sph.so; the product has no fallback, so a missing library is a failure."""
    import __graft_entry__ as ge
    import dualsphysics_b200 as dsb
    if not os.path.exists(dsb.LIB_PATH):
        ge.build()
    return dsb.load_library()
