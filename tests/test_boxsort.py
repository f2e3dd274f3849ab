"""Row N4 (per-box read ordering and duplicate marking): the C restatement and the device program (host build) against the
reference's own bubblesortReads / markDuplicates (src/reporter.cpp:932-972) -- committed dumps and, when oracle/_ref is
built, live.  The GPU part is tests/test_gpu_boxsort.py."""
import os

import numpy as np
import pytest

from oracle import boxsortoracle as B, refhook
from tests import boxsort_helpers as H
from tests.emul import boxsort_emul
from tests.golden import make_boxsort_golden as G

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def lib():
    return boxsort_emul.build()


@pytest.mark.parametrize("name", list(G.CASES))
def test_oracle_and_device_program_match_reference_dumps(lib, name):
    z = np.load(os.path.join(GOLD, "boxsort_%s.npz" % name))
    reads, box_off, idx = B.make_boxes(**G.CASES[name])
    o_idx, o_unique = B.sort_boxes(reads, box_off, idx)
    assert np.array_equal(o_idx, z["sorted_idx"]) and np.array_equal(o_unique, z["unique"]), "oracle"
    for warp_formulation in (False, True):      # the literal replay, and the kernel's chunked weak-prefix-minimum form
        order, non = boxsort_emul.sort_boxes(lib, reads[idx], box_off, warp_formulation)
        p_idx, p_unique = H.apply(reads, box_off, idx, order, non)
        assert np.array_equal(p_idx, z["sorted_idx"]) and np.array_equal(p_unique, z["unique"]), ("device program", warp_formulation)
    # the exchange sort is not a stable sort: a stable sort of the same keys gives another order
    stable = idx.copy()
    for b in range(len(box_off) - 1):
        a, e = int(box_off[b]), int(box_off[b + 1])
        k = reads[idx[a:e]]
        stable[a:e] = idx[a:e][np.lexsort((k["bp"], k["nt_size"], k["indel_size"], k["bp_right"], k["bp_left"]))]
    assert not np.array_equal(stable, z["sorted_idx"])


@pytest.mark.skipif(not refhook.available(), reason="oracle/_ref not built")
def test_oracle_matches_live_reference():
    reads, box_off, idx = B.make_boxes(150, 77, max_reads=60, heavy=2)
    r_idx, r_unique = B.reference_sort_boxes(reads, box_off, idx)
    o_idx, o_unique = B.sort_boxes(reads, box_off, idx)
    assert np.array_equal(o_idx, r_idx) and np.array_equal(o_unique, r_unique)
