"""GPU parity of the discordant-read-pair clustering: pg_rp_modify / pg_rp_summarize
through the C ABI against the reference's own dumps (tests/golden/rp_*.npz) and against the oracle on fresh seeded sets,
including groups larger than one CTA's shared memory."""
import os

import numpy as np
import pytest

from oracle import rporacle
from pindel_b200 import api, rp, synth
from tests.golden import make_rp_golden as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pg():
    p = api.PatternGrowth()
    yield p
    p.close()


@pytest.mark.parametrize("name", list(G.CASES))
def test_matches_reference_dumps(pg, name):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "rp_%s.npz" % name))
    t = z["table"]
    rows = G.rows_from_table(t)
    before = pg.rp_launches
    got = pg.rp_summarize(rows, pg.rp_modify(rows))
    assert pg.rp_launches > before or rows.size < 5
    for col, key in ((7, "pos_a"), (8, "pos_a1"), (9, "pos_b"), (10, "pos_b1"), (11, "visited"), (12, "report"), (13, "n_identical")):
        assert np.array_equal(got[key].astype(np.int64), t[:, col]), key
    # Tags: rows merged away name the row that holds their Tags in the reference
    tags, p = [], 0
    for c in z["tag_counts"]:
        tags.append(set(z["tag_values"][p:p + c].tolist()))
        p += c
    holder = {}
    for i, tg in enumerate(tags):
        for x in tg:
            holder[x] = i
    for i in range(rows.size):
        assert holder[int(t[i, 0])] == int(got["merged_into"][i]), i


@pytest.mark.parametrize("n_events,seed,span,support", [(300, 31, 2_000_000, 50), (2500, 32, 40_000_000, 40), (60, 33, 200_000, 400)])
def test_matches_oracle_stage_by_stage(pg, n_events, seed, span, support):
    """(60, 33): 400 pairs per event -> long chains of overlapping intervals.  (2500, 32): ~60 000 rows, groups of ~15 000:
    larger than the 12 800 rows of intervals a CTA keeps in shared memory."""
    rows = synth.make_discordant_pairs(n_events, seed, span=span, background=n_events, max_support=support)
    rows = rows[rp.compare2rp_order(rows)]
    mod = pg.rp_modify(rows)
    exp_mod = rporacle.modify(rows)
    assert mod.tobytes() == exp_mod.tobytes()
    assert pg.rp_summarize(rows, mod).tobytes() == rporacle.summarize(rows, exp_mod).tobytes()


def test_small_and_empty_inputs(pg):
    rows = synth.make_discordant_pairs(1, 5, max_support=2, both_mates=0.0)[:3]
    out = pg.rp_summarize(rows, pg.rp_modify(rows))
    assert out.tobytes() == rporacle.cluster(rows).tobytes() and not out["report"].any()
    assert pg.rp_modify(rows[:0]).size == 0
    odd = rows.copy()
    odd["da"] = ord("k")          # RP_READ's default strand (src/pindel.h:221): no interval test can pass
    assert pg.rp_modify(odd).tobytes() == rporacle.modify(odd).tobytes()


@pytest.mark.parametrize("name", list(G.INTER_CASES))
def test_interchr_matches_reference_dumps(pg, name):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "rp_%s.npz" % name))
    t = z["table"]
    rows = G.rows_from_table(t)
    before = pg.rp_launches
    _, got = pg.rp_cluster_interchr(rows, z["chr_a"], z["chr_b"])
    assert pg.rp_launches > before
    for col, key in ((7, "pos_a"), (8, "pos_a1"), (9, "pos_b"), (10, "pos_b1"), (11, "visited"), (12, "report"), (13, "n_identical")):
        assert np.array_equal(got[key].astype(np.int64), t[:, col]), key


def test_interchr_matches_oracle_stage_by_stage(pg):
    rows, ca, cb = synth.make_interchr_pairs(400, 61, n_chr=4, max_support=60, background=300)
    order = rp.compare2rp_order(rows)
    rows, ca, cb = rows[order], ca[order], cb[order]
    mod, fin = pg.rp_cluster_interchr(rows, ca, cb)
    omod, ofin = rporacle.cluster_interchr(rows, ca, cb)
    assert mod.tobytes() == omod.tobytes() and fin.tobytes() == ofin.tobytes()
    one = pg.rp_cluster_interchr(rows[:1], ca[:1], cb[:1])
    assert one[1].tobytes() == rporacle.cluster_interchr(rows[:1], ca[:1], cb[:1])[1].tobytes()
