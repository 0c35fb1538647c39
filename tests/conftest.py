import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(autouse=True)
def _gpu_watchdog(request):
    """A GPU test that is stuck inside a CUDA call cannot be interrupted by pytest-timeout's signal method (the handler
    only runs once the C call returns), and a process that never exits holds the GPU box until it is killed from outside.
    So every gpu-marked test also arms a hard watchdog: stacks are dumped and the process exits if one test runs for
    25 minutes (far beyond the slowest test; the per-test pytest timeouts stay the normal mechanism)."""
    if request.node.get_closest_marker("gpu") is None:
        yield
        return
    import faulthandler
    faulthandler.dump_traceback_later(1500, exit=True)
    try:
        yield
    finally:
        faulthandler.cancel_dump_traceback_later()


@pytest.fixture(scope="session", autouse=True)
def _scratch_cwd(tmp_path_factory):
    # the reference's forward() opens layer_outputs_cpu.txt in the cwd on every call (model.cpp:42)
    d = tmp_path_factory.mktemp("cwd")
    old = os.getcwd()
    os.chdir(d)
    yield
    os.chdir(old)


@pytest.fixture(scope="session")
def port():
    from oracle import loader
    loader.build("port")
    return loader.Port()


@pytest.fixture(scope="session")
def ref():
    from oracle import loader
    if not loader.have_ref():
        if os.path.isdir("/root/reference"):
            loader.build("ref")
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return loader.Ref()


@pytest.fixture(scope="session")
def golden_ops():
    return dict(np.load(os.path.join(GOLDEN, "ops_ref.npz")))


@pytest.fixture(scope="session")
def golden_models():
    return dict(np.load(os.path.join(GOLDEN, "models_ref.npz")))


def oracle_shape(ms):
    """simplellminference_b200.ModelShape -> oracle.loader.Shape"""
    from oracle import loader
    return loader.Shape(ms.vocab, ms.head_dim, ms.hidden, ms.kv_hidden, ms.inter, ms.max_len, ms.layers, ms.heads,
                        ms.kv_heads, ms.eps, ms.theta)
