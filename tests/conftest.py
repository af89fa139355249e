import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import __graft_entry__ as graft  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "tiny")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    """The host binding; building is the driver's job (graft.build()), but build lazily if the
    shared library is absent so a fresh checkout can run the CPU suite."""
    p = graft.load_package()
    if not os.path.exists(p.LIB_PATH):
        graft.build()
    p.lib()
    return p


@pytest.fixture(scope="session")
def orc():
    return graft.load_oracle()


@pytest.fixture(scope="session")
def golden():
    z = np.load(os.path.join(GOLDEN, "cases.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_onnx():
    return os.path.join(GOLDEN, "model.onnx")


@pytest.fixture(scope="session")
def model_cache(tmp_path_factory):
    """Directory for ONNX files exported on the fly (mini / small / base archs)."""
    d = os.environ.get("GLC_MODEL_CACHE") or str(tmp_path_factory.mktemp("glc_models"))
    os.makedirs(d, exist_ok=True)
    return d
