"""Searcher::loadRegion / getCorrespondingSearchWindowClusters of the C++ mirror (include/pindel_pg.hpp) on the GPU, against the
oracle's position sweep.  (Its host logic is also checked on the CPU against the emulated ABI, tests/test_cpp_host_mirror.py.)"""
import pytest

from tests import test_cpp_host_mirror as M

pytestmark = pytest.mark.gpu


def test_bd_mirror_matches_oracle(tmp_path):
    exe = M.build_exe(M.BD_SRC, M.BD_EXE)
    for seed in (1, 2, 3, 6):
        M.bd_mirror_case(exe, tmp_path, seed)
