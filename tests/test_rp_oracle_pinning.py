"""Row N2 (discordant-pair clustering): the C restatement oracle/rp_oracle.c against the reference's own ModifyRP /
Summarize (src/bddata.cpp:385-431, 470-575) -- live through oracle/_ref/libpindel_ref.so when it is built here, and against
the committed dumps of the same call (tests/golden/rp_*.npz, made by tests/golden/make_rp_golden.py)."""
import os

import numpy as np
import pytest

from oracle import refhook, rporacle
from pindel_b200 import synth
from tests.golden import make_rp_golden as G


def check(rows_sorted, table, tags):
    got = rporacle.cluster(rows_sorted)
    for col, key in ((7, "pos_a"), (8, "pos_a1"), (9, "pos_b"), (10, "pos_b1"), (11, "visited"), (12, "report"), (13, "n_identical")):
        bad = np.nonzero(got[key].astype(np.int64) != table[:, col])[0]
        assert bad.size == 0, "%s differs at %d rows, first %d: %d vs %d" % (key, bad.size, bad[0], got[key][bad[0]], table[bad[0], col])
    # Tags: a row that was merged away gave its Tags to merged_into
    members = {}
    for i, m in enumerate(got["merged_into"]):
        members.setdefault(int(m), []).append(int(table[i, 0]))
    for i in range(len(tags)):
        exp = sorted(tags[i])
        mine = sorted(members.get(i, [])) if int(got["merged_into"][i]) == i else []
        assert mine == exp, (i, mine, exp)


@pytest.mark.parametrize("name", list(G.CASES))
def test_port_matches_committed_reference_dumps(name):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "rp_%s.npz" % name))
    tags, p = [], 0
    for c in z["tag_counts"]:
        tags.append(z["tag_values"][p:p + c].tolist())
        p += c
    check(G.rows_from_table(z["table"]), z["table"], tags)


@pytest.mark.skipif(not refhook.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [11, 12, 13])
def test_port_matches_live_reference(seed):
    rows = synth.make_discordant_pairs(120, seed, span=3_000_000 if seed == 13 else 60_000_000, background=300)
    table, tags, _ = refhook.rp_cluster(rows)
    check(G.rows_from_table(table), table, tags)


def test_fewer_than_five_rows_report_nothing():
    rows = synth.make_discordant_pairs(1, 5, max_support=2, both_mates=0.0)[:4]
    out = rporacle.cluster(rows)
    assert not out["report"].any()
