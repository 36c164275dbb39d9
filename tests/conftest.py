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


class _OrderedContext:
    """The context's streams are non-blocking: nothing orders them behind torch's stream, on which a test has just filled its
    device tensors (torch.zeros / ones / full are asynchronous kernels).  A production caller orders its streams itself; the
    tests simply let torch finish before every call into the library -- without this a prefill could land AFTER the library's
    own memset and show up as stale values in macroblocks the kernel does not write (seen once in 200 runs)."""

    def __init__(self, inner):
        object.__setattr__(self, "_inner", inner)

    def __getattr__(self, name):
        attr = getattr(self._inner, name)
        if not callable(attr) or name.startswith("_") or name in ("close", "sync", "torch_stream", "pinned_empty"):
            return attr

        def call(*args, **kwargs):
            import torch
            if torch.cuda.is_available() and torch.cuda.is_initialized():
                torch.cuda.synchronize()
            return attr(*args, **kwargs)
        return call


@pytest.fixture(scope="session")
def ctx(pkg):
    """a device context; GPU tests fail (not skip) when the CUDA library cannot open a device"""
    c = pkg.Context(0)
    yield _OrderedContext(c)
    c.close()
