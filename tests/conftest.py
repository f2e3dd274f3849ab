import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import pgoracle
    return pgoracle.load()


@pytest.fixture(scope="session", autouse=True)
def native_artefacts():
    """The suite needs the in-tree native builds (the product library for the ABI tests, the oracle port, the emulator);
    build whatever is missing or stale once per session -- the same thing __graft_entry__.build() does."""
    from pindel_b200 import build as B
    B.build_cuda()
    B.build_oracle()
    yield
