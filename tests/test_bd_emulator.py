"""The device formulation of the BreakDancer mask (pg_bd.cuh: two binary searches per read over the window-start array and the
prefix-maximum drop array, the cluster built per read) in its host build, against the oracle's position-by-position sweep
(pinned to the reference by test_bd_oracle_pinning.py) and, where the reference library is built, against the reference live."""
import os

import numpy as np
import pytest

from oracle import bdoracle, refhook
from tests import bd_helpers as B
from tests.emul import bd_emul


@pytest.fixture(scope="module")
def lib():
    return bd_emul.build()


def case(seed, dense):
    rng = np.random.default_rng(500 + seed)
    chr_id = int(rng.integers(0, 3))
    ws = int(rng.choice([B.SPACER, B.SPACER + 500, B.SPACER + 40_000, 1000, 0]))
    window = (ws, ws + int(rng.choice([5_000, 20_000, 60_000])))
    ev = B.random_events(rng, int(rng.choice([0, 1, 3, 40, 200])) if not dense else 200, chr_id, window, dense)
    other = B.random_events(rng, 20, (chr_id + 1) % 3, window)
    ev = np.concatenate([ev, other])
    return chr_id, window, ev, B.queries(rng, window, 600)


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


@pytest.mark.skipif(not refhook.available(), reason="oracle/_ref/libpindel_ref.so not built")
def test_closed_form_equals_reference_live(lib, tmp_path):
    rng = np.random.default_rng(5)
    fa = os.path.join(tmp_path, "g.fa")
    refhook.write_fasta(fa, [(n, rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), l)) for n, l in zip(B.NAMES, B.LENS)])
    hook = refhook.RefHook(fa, [], workdir=str(tmp_path))
    for seed in range(30, 40):
        chr_id, window, ev, q = case(seed, seed % 2 == 0)
        exp = bdoracle.reference_clusters(hook, ev, chr_id, window[0], window[1], q)
        ev_sorted = np.array(bdoracle.sort_on_first(bdoracle.normalise(ev), B.NAMES), dtype=np.int64).reshape(-1, 6)
        got = bd_emul.clusters(lib, ev_sorted, B.NAMES, chr_id, window[0], window[1], q)
        we = window[1] + 3000
        for k, (e, g) in enumerate(zip(exp, got)):
            if q[k] >= we:
                continue
            assert e == g, (seed, k, int(q[k]), e, g)
