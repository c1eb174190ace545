import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _product_library_is_built():
    """Tests that spawn worker processes import the package on their own: make sure the in-tree library exists before
    anything runs (a fresh checkout has no build artefacts; nvcc cross-compiles without a GPU)."""
    import snark_challenge_prover_reference_b200 as b
    if not os.path.exists(b.LIB_PATH):
        from snark_challenge_prover_reference_b200 import build as _b
        _b.build()


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (C restatement of the reference). Test infrastructure only."""
    import util
    return util.load_oracle()


@pytest.fixture(scope="session")
def b200():
    """The product library (must already be built in-tree; building is __graft_entry__.build()'s job)."""
    import snark_challenge_prover_reference_b200 as b
    if not os.path.exists(b.LIB_PATH):
        from snark_challenge_prover_reference_b200 import build as _b
        _b.build()
    b.lib()
    return b
