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


KINDS = {0: "D/TD", 1: "SI", 2: "DI", 3: "INV", 4: "INV_NT"}


def kind_case(kind, seed, heavy=0):
    if kind <= 1:
        reads, box_off, idx = B.make_boxes(60, seed, max_reads=50, heavy=heavy)
        rng = np.random.default_rng(seed)
        reads["indel_size"] = np.where(rng.random(reads.size) < 0.3, reads["indel_size"] + 1, reads["indel_size"])     # SI events: (BPLeft, IndelSize)
        return reads, box_off, idx
    return B.make_boxes_kind(kind, 60, seed, max_reads=50, heavy=heavy)


@pytest.mark.parametrize("kind", list(KINDS))
@pytest.mark.parametrize("seed,heavy", [(1, 0), (2, 0), (3, 2)])
def test_device_program_equals_the_restated_reporter_loops(lib, kind, seed, heavy):
    """every reporter's ordering, duplicate marks and event cut: the literal device form and the kernel's formulation (host
    build) against oracle/boxsort_oracle.c's restatement of the reference's loops"""
    reads, box_off, idx = kind_case(kind, 100 * kind + seed, heavy)
    o_idx, o_unique, o_event = B.sort_boxes_kind(kind, reads, box_off, idx)
    for warp_formulation in (False, True):
        order, non, ev = boxsort_emul.sort_boxes_kind(lib, kind, reads[idx], box_off, warp_formulation)
        p_idx, p_unique = H.apply(reads, box_off, idx, order, non)
        assert np.array_equal(p_idx, o_idx), (KINDS[kind], "order", warp_formulation)
        assert np.array_equal(p_unique, o_unique), (KINDS[kind], "duplicate marks", warp_formulation)
        assert np.array_equal(ev, o_event), (KINDS[kind], "event cut", warp_formulation)
    # the inputs reach the corners: ties that a stable sort would order differently, duplicates, several events per box
    n_boxes = len(box_off) - 1
    assert (o_event & 2).sum() > n_boxes and (o_unique == 0).sum() > reads.size // 50
    if kind >= 2:
        assert not np.array_equal(o_idx, idx)


def test_kind_d_is_the_plain_call(lib):
    reads, box_off, idx = B.make_boxes(40, 9, max_reads=30)
    a = boxsort_emul.sort_boxes(lib, reads[idx], box_off, True)
    b = boxsort_emul.sort_boxes_kind(lib, 0, reads[idx], box_off, True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
