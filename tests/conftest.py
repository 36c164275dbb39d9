"""pytest configuration: the `gpu` marker, package loading, shared fixtures."""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_package():
    """import x264-dsp_b200/ (hyphenated directory) as module `x264dsp_b200`"""
    if "x264dsp_b200" in sys.modules:
        return sys.modules["x264dsp_b200"]
    path = os.path.join(ROOT, "x264-dsp_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location("x264dsp_b200", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["x264dsp_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pkg():
    mod = load_package()
    mod.build()
    return mod


@pytest.fixture(scope="session")
def ctx(pkg):
    """a device context; GPU tests fail (not skip) when the CUDA library cannot open a device"""
    c = pkg.Context(0)
    yield c
    c.close()
