import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def deck_dir(tmp_path_factory):
    """Generated decks written to a temp dir (ref_harness wants files)."""
    from minimc_b200 import decks
    d = tmp_path_factory.mktemp("decks")
    return d
