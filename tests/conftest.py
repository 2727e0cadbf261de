import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_library():
    """libb200tts.so is a build artefact (git-ignored): a fresh checkout compiles it once (nvcc cross-compiles sm_100a
    without a GPU) so that the ABI tests do not depend on a previous `python -m larynx_b200.build`."""
    from larynx_b200 import _lib, build

    if not _lib.lib_path().exists():
        build.build()
    return _lib.load()
