"""GPU parity of row N4 (per-box read ordering + duplicate marking): pg_box_sort through the C ABI against the reference's
own dumps (tests/golden/boxsort_*.npz) and the oracle on a fresh, larger set."""
import os

import numpy as np
import pytest

from oracle import boxsortoracle as B
from pindel_b200 import api
from tests import boxsort_helpers as H
from tests.golden import make_boxsort_golden as G

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def pg():
    p = api.PatternGrowth()
    yield p
    p.close()


@pytest.mark.parametrize("name", list(G.CASES))
def test_matches_reference_dumps(pg, name):
    z = np.load(os.path.join(GOLD, "boxsort_%s.npz" % name))
    reads, box_off, idx = B.make_boxes(**G.CASES[name])
    before = pg.box_sort_launches
    order, non = pg.box_sort(reads[idx], box_off)
    assert pg.box_sort_launches == before + 1
    p_idx, p_unique = H.apply(reads, box_off, idx, order, non)
    assert np.array_equal(p_idx, z["sorted_idx"]) and np.array_equal(p_unique, z["unique"])


def test_matches_oracle_on_many_boxes(pg):
    reads, box_off, idx = B.make_boxes(20_000, 5, max_reads=30, heavy=3)
    order, non = pg.box_sort(reads[idx], box_off)
    p_idx, p_unique = H.apply(reads, box_off, idx, order, non)
    o_idx, o_unique = B.sort_boxes(reads, box_off, idx)
    assert np.array_equal(p_idx, o_idx) and np.array_equal(p_unique, o_unique)


def test_empty_and_single_read_boxes(pg):
    reads, box_off, idx = B.make_boxes(5, 9)
    box_off = np.array([0, 0, 1, 1, box_off[-1]], dtype=np.uint32)      # empty, one read, empty, the rest
    order, non = pg.box_sort(reads[idx], box_off)
    p_idx, p_unique = H.apply(reads, box_off, idx, order, non)
    o_idx, o_unique = B.sort_boxes(reads, box_off, idx)
    assert np.array_equal(p_idx, o_idx) and np.array_equal(p_unique, o_unique)
    o, n = pg.box_sort(reads[:0], np.array([0], dtype=np.uint32))
    assert o.size == 0 and n.size == 0


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 4])
def test_every_reporters_loops_match_oracle(pg, kind):
    """pg_box_sort_kind: ordering, duplicate marks and event cut of the D/TD, SI, DI, INV and INV_NT reporters on a few thousand
    boxes (three of them with hundreds of reads) against oracle/boxsort_oracle.c's restatement of the reference's loops"""
    if kind <= 1:
        reads, box_off, idx = B.make_boxes(3000, 40 + kind, max_reads=40, heavy=3)
    else:
        reads, box_off, idx = B.make_boxes_kind(kind, 3000, 40 + kind, max_reads=40, heavy=3)
    order, non, ev = pg.box_sort_kind(kind, reads[idx], box_off)
    p_idx, p_unique = H.apply(reads, box_off, idx, order, non)
    o_idx, o_unique, o_event = B.sort_boxes_kind(kind, reads, box_off, idx)
    assert np.array_equal(p_idx, o_idx), "order"
    assert np.array_equal(p_unique, o_unique), "duplicate marks"
    assert np.array_equal(ev, o_event), "event cut"
    assert (o_event & 2).sum() > 3000 and (o_unique == 0).sum() > 1000


def test_unknown_kind_is_an_error(pg):
    reads, box_off, idx = B.make_boxes(3, 1)
    with pytest.raises(api.PgError):
        pg.box_sort_kind(7, reads[idx], box_off)
