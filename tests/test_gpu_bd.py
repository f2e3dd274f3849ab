"""GPU tests of the BreakDancer mask on the device (last part): pg_bd_load_region +
pg_bd_clusters against the oracle's position-by-position sweep (oracle/bdoracle.py, pinned to the reference's own
BDData::loadRegion / getCorrespondingSearchWindowCluster by tests/test_bd_oracle_pinning.py), and pg_far_end taking every
read's cluster from the loaded region against the same call with host-supplied clusters and against the search oracle."""
import numpy as np
import pytest

from oracle import bdoracle
from pindel_b200 import api, runs as R, synth
from tests import bd_helpers as B
from tests import helpers as H
from tests.emul.bd_emul import name_ranks

pytestmark = pytest.mark.gpu


def sorted_events(ev, names):
    return np.array(bdoracle.sort_on_first(bdoracle.normalise(ev), names), dtype=np.int64).reshape(-1, 6)


def as_lists(off, reg):
    return [[(int(r["chr"]), int(r["start"]), int(r["end"])) for r in reg[off[i]:off[i + 1]]] for i in range(off.size - 1)]


@pytest.fixture(scope="module")
def pg3():
    rng = np.random.default_rng(77)
    pg = api.PatternGrowth()
    assert B.SPACER == synth.SPACER
    pg.set_reference([synth.pad_reference(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), l)) for l in B.LENS])
    yield pg
    pg.close()


@pytest.mark.parametrize("seed", range(1, 13))
def test_clusters_equal_oracle(pg3, seed):
    rng = np.random.default_rng(900 + seed)
    chr_id = int(rng.integers(0, 3))
    ws = int(rng.choice([B.SPACER, B.SPACER + 500, B.SPACER + 40_000, 1000, 0]))
    window = (ws, ws + int(rng.choice([5_000, 20_000, 60_000])))
    dense = seed % 3 == 0
    ev = B.random_events(rng, 200 if dense else int(rng.choice([0, 1, 3, 40, 200])), chr_id, window, dense)
    ev = np.concatenate([ev, B.random_events(rng, 20, (chr_id + 1) % 3, window)])
    q = B.queries(rng, window, 5000)
    exp = bdoracle.clusters(ev, B.NAMES, chr_id, window[0], window[1], q)
    pg3.bd_load_region(sorted_events(ev, B.NAMES), name_ranks(B.NAMES), chr_id, window[0], window[1])
    got = as_lists(*pg3.bd_clusters(q))
    for k, (e, g) in enumerate(zip(exp, got)):
        if e is None:
            assert g == []
            continue
        assert e == g, (seed, k, int(q[k]), e, g)
    pg3.bd_unload()


@pytest.mark.parametrize("k", range(len(B.load_golden())))
def test_clusters_equal_reference_outputs(pg3, k):
    """the committed outputs of the reference's own loadRegion / getCorrespondingSearchWindowCluster (tests/golden/bd_clusters.npz)"""
    seed, chr_id, window, ev, q, exp = B.load_golden()[k]
    pg3.bd_load_region(sorted_events(ev, B.NAMES), name_ranks(B.NAMES), chr_id, window[0], window[1])
    got = as_lists(*pg3.bd_clusters(q))
    for i, (e, g) in enumerate(zip(exp, got)):
        assert e == g, (seed, i, int(q[i]), e, g)
    pg3.bd_unload()


def test_load_region_rejects_unsorted_events_and_lookup_without_region(pg3):
    ev = np.array([[0, 121000, 121000, 1, 140000, 140000], [0, 120000, 120300, 0, 150000, 150500]], dtype=np.int64)
    with pytest.raises(api.PgError):
        pg3.bd_load_region(ev, name_ranks(B.NAMES), 0, 110000, 130000)
    with pytest.raises(api.PgError):
        pg3.bd_clusters(np.array([120100], dtype=np.uint32))
    pg3.bd_load_region(ev[::-1], name_ranks(B.NAMES), 0, 110000, 130000)
    off, reg = pg3.bd_clusters(np.array([], dtype=np.uint32))
    assert off.tolist() == [0] and reg.size == 0
    pg3.bd_unload()


def close_seam(batch, close):
    """host side of the GetCloseEnd / SearchFarEnds seam: flipped sequences and the close end SearchFarEnds starts from"""
    seqs = batch.seq.copy()
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    ci = np.zeros(batch.n, dtype=R.CLOSEINFO_DT)
    for i, r in enumerate(close):
        if r["flip"]:
            seqs[i] = comp[seqs[i][::-1]]
        if len(r["close"]):
            ci["last_abs_loc"][i] = r["close"][-1][1]
            ci["max_len"][i] = r["close"][-1][0]
    return seqs, ci


def test_far_end_takes_clusters_from_loaded_region():
    """SearchFarEnd with BreakDancer clusters (src/pindel.cpp:1006-1020): far ends beyond the distance ranges are found only
    through the cluster of the read's close end.  Device lookup == host-supplied clusters == the search oracle."""
    ref = synth.make_reference(700_000, 20261017)
    padded = synth.pad_reference(ref)
    batch = synth.make_reads(ref, 6000, 100, seed=77, max_del=40000, mix=(0.9, 0.05, 0.0, 0.05))
    kw = H.batch_kwargs(batch)
    wide = api.PatternGrowth(max_range_index=6)
    wide.set_reference([padded])
    truth = R.per_read(*wide.search(**kw), kw["seq_blob"], kw["seq_off"])
    wide.close()

    pg = api.PatternGrowth(max_range_index=1)
    pg.set_reference([padded])
    close = R.per_read(*pg.get_close_end(**kw))
    seqs, ci = close_seam(batch, close)
    fkw = dict(seq_blob=seqs.reshape(-1), seq_off=kw["seq_off"], close_info=ci, matched_d=kw["matched_d"],
               matched_rel_pos=kw["matched_rel_pos"], insert_size=kw["insert_size"])
    plain = R.per_read(*pg.search_far_ends(**fkw))

    # events: (close end, far end found with wide ranges) of every third read that has both, plus unrelated ones
    rng = np.random.default_rng(5)
    ev = []
    for i, t in enumerate(truth):
        if i % 3 == 0 and ci["max_len"][i] and len(t["far"]):
            a = int(ci["last_abs_loc"][i]) + int(rng.integers(-150, 150))
            b = int(t["far"][-1][1]) + int(rng.integers(-100, 100))
            ev.append((0, a, a + int(rng.integers(0, 60)), 0, b, b + int(rng.integers(0, 60))))
            ev.append((0, b, b, 0, a, a))
    ev = np.array(ev, dtype=np.int64).reshape(-1, 6)
    window = (synth.SPACER + 50_000, synth.SPACER + 600_000)
    names = ["chr1"]
    exp_clusters = bdoracle.clusters(ev, names, 0, window[0], window[1], ci["last_abs_loc"])
    roff = np.zeros(batch.n + 1, dtype=np.uint32)
    flat = []
    for i, c in enumerate(exp_clusters):
        if ci["max_len"][i] and c:
            flat.extend(c)
        roff[i + 1] = len(flat)
    rg = np.zeros(len(flat), dtype=R.REGION_DT)
    a = np.array(flat, dtype=np.int64).reshape(-1, 3)
    rg["chr"], rg["start"], rg["end"] = a[:, 0], a[:, 1], a[:, 2]
    assert rg.size > 100

    host_side = R.per_read(*pg.search_far_ends(regions=rg, region_off=roff, **fkw))
    pg.bd_load_region(sorted_events(ev, names), name_ranks(names), 0, window[0], window[1])
    before = pg.bd_lookups
    out_d, runs_d = pg.search_far_ends(**fkw)
    device_side = R.per_read(out_d, runs_d)
    assert pg.bd_lookups == before + batch.n
    differ = 0
    for i in range(batch.n):
        assert np.array_equal(host_side[i]["far"], device_side[i]["far"]), i
        differ += not np.array_equal(plain[i]["far"], device_side[i]["far"])
    assert differ >= 10, differ           # the clusters matter: far ends out of reach of max_range_index=1

    # and against the search oracle given the oracle's clusters
    exp = H.oracle_results([padded], batch, regions=a.astype(np.uint32), region_off=roff, max_range_index=1)
    for i in range(batch.n):
        assert np.array_equal(exp[i]["far"], device_side[i]["far"]), i

    # a batch that brings its own clusters ignores the loaded region; unloading restores the plain search
    again = R.per_read(*pg.search_far_ends(regions=rg[:1], region_off=np.zeros(batch.n + 1, np.uint32), **fkw))
    pg.bd_unload()
    after = R.per_read(*pg.search_far_ends(**fkw))
    for i in range(batch.n):
        assert np.array_equal(plain[i]["far"], again[i]["far"]) and np.array_equal(plain[i]["far"], after[i]["far"]), i
    pg.close()
