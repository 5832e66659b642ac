import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the @pytest.mark.gpu tests are skipped (a plain `pytest tests` on a CPU box stays green);
    on a GPU box nothing is skipped -- and the product has no CPU path, so a missing extension fails loudly there."""
    gpu_items = [item for item in items if "gpu" in item.keywords]
    if not gpu_items:
        return
    try:
        from minimc_b200 import capi
        have_gpu = capi.load().mmc_device_count() > 0
    except OSError:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (minimc_b200 has no CPU transport path)")
    for item in gpu_items:
        item.add_marker(skip)


@pytest.fixture(scope="session")
def deck_dir(tmp_path_factory):
    """Generated decks written to a temp dir (ref_harness wants files)."""
    from minimc_b200 import decks
    d = tmp_path_factory.mktemp("decks")
    return d
