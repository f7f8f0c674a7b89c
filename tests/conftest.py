import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build the CUDA library and the CPU checkers once per session (no-ops when up to date)."""
    from rawhash_b200 import build
    build.build_lib()
    build.build_cli()
    build.stage_models()
    build.build_oracle()
    return True
