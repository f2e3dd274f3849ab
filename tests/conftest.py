import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    """The whole-program scenarios (the reference host on the library, tens of seconds each) run after the kernel-level tests:
    under `-x` a scenario that fails must not hide the parity tests that would follow it alphabetically."""
    late = [it for it in items if "test_gpu_dropin" in it.nodeid]
    if late:
        items[:] = [it for it in items if "test_gpu_dropin" not in it.nodeid] + late


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
