import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ref_svox2():
    """The UNMODIFIED reference package (pure-PyTorch + its CUDA extension built for sm_100),
    installed by oracle/build_ref.sh into oracle/_ref (git-ignored, travels to the GPU box)."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "svox2")):
        pytest.skip("oracle/_ref not built (run oracle/build_ref.sh in the build container)")
    sys.path.insert(0, ref_dir)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import svox2
    return svox2
