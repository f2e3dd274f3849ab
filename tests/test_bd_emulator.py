"""The device formulation of the BreakDancer mask (pg_bd.cuh: two binary searches per read over the window-start array and the
prefix-maximum drop array, the cluster built per read) in its host build, against the oracle's position-by-position sweep
(pinned to the reference by test_bd_oracle_pinning.py) and against the committed outputs of the reference itself."""
import numpy as np
import pytest

from oracle import bdoracle
from tests import bd_helpers as B
from tests.emul import bd_emul


@pytest.fixture(scope="module")
def lib():
    return bd_emul.build()


case = B.case


@pytest.mark.parametrize("seed,dense", [(s, s % 3 == 0) for s in range(1, 25)])
def test_closed_form_equals_sweep(lib, seed, dense):
    chr_id, window, ev, q = case(seed, dense)
    exp = bdoracle.clusters(ev, B.NAMES, chr_id, window[0], window[1], q)
    ev_sorted = np.array(bdoracle.sort_on_first(bdoracle.normalise(ev), B.NAMES), dtype=np.int64).reshape(-1, 6)
    got = bd_emul.clusters(lib, ev_sorted, B.NAMES, chr_id, window[0], window[1], q)
    for k, (e, g) in enumerate(zip(exp, got)):
        if e is None:
            assert g == []                    # the reference reads an entry its sweep never wrote: defined as "no cluster" here
            continue
        assert e == g, (seed, k, int(q[k]), e, g)


def test_unnormalised_positions_and_unsorted_block(lib):
    """pos > pos2 is ordered as the reference's constructor does; a block that is not ascending is an error, not a guess"""
    ev = np.array([[0, 120300, 120000, 0, 150500, 150000], [0, 121000, 121000, 1, 140000, 140000]], dtype=np.int64)
    exp = bdoracle.clusters(ev, B.NAMES, 0, 110000, 130000, [120100, 120700, 121100])
    got = bd_emul.clusters(lib, ev, B.NAMES, 0, 110000, 130000, [120100, 120700, 121100])
    assert exp == got and any(got)
    with pytest.raises(AssertionError):
        bd_emul.clusters(lib, ev[::-1], B.NAMES, 0, 110000, 130000, [120100])


@pytest.mark.parametrize("k", range(len(B.load_golden())))
def test_closed_form_equals_reference_outputs(lib, k):
    """the committed outputs of the reference's own loadRegion / getCorrespondingSearchWindowCluster (tests/golden/bd_clusters.npz)"""
    seed, chr_id, window, ev, q, exp = B.load_golden()[k]
    ev_sorted = np.array(bdoracle.sort_on_first(bdoracle.normalise(ev), B.NAMES), dtype=np.int64).reshape(-1, 6)
    got = bd_emul.clusters(lib, ev_sorted, B.NAMES, chr_id, window[0], window[1], q)
    for i, (e, g) in enumerate(zip(exp, got)):
        assert e == g, (seed, i, int(q[i]), e, g)
