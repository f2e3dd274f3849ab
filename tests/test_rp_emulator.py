"""The read-pair clustering program (pindel_b200/csrc/pg_rp.cuh: the predicates pg_rp_modify_kernel / pg_rp_group_kernel
run) compiled for the host and driven in the kernels' own formulation -- find first applicable row, update, rescan; first
earlier identical row -- against the reference's dumps and the oracle.  No GPU."""
import os

import numpy as np
import pytest

from oracle import rporacle
from pindel_b200 import synth
from tests.emul import rp_emul
from tests.golden import make_rp_golden as G


@pytest.fixture(scope="module")
def lib():
    return rp_emul.build()


@pytest.mark.parametrize("name", list(G.CASES))
def test_program_matches_reference_dumps(lib, name):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "rp_%s.npz" % name))
    t = z["table"]
    _, got = rp_emul.cluster(lib, G.rows_from_table(t))
    for col, key in ((7, "pos_a"), (8, "pos_a1"), (9, "pos_b"), (10, "pos_b1"), (11, "visited"), (12, "report"), (13, "n_identical")):
        assert np.array_equal(got[key].astype(np.int64), t[:, col]), key


@pytest.mark.parametrize("seed", [21, 22])
def test_program_matches_oracle_stage_by_stage(lib, seed):
    rows = synth.make_discordant_pairs(300, seed, span=2_000_000, background=200, max_support=50)
    order = np.lexsort((-rows["pos_b"].astype(np.int64), -rows["pos_a"].astype(np.int64)))   # Compare2RP: descending
    rows = rows[order]
    mod, fin = rp_emul.cluster(lib, rows)
    exp_mod = rporacle.modify(rows)
    assert mod.tobytes() == exp_mod.tobytes()
    assert fin.tobytes() == rporacle.summarize(rows, exp_mod).tobytes()


@pytest.mark.parametrize("name", list(G.INTER_CASES))
def test_interchr_program_and_oracle_match_reference_dumps(lib, name):
    """ModifyRP_InterChr + Summarize_InterChr: the kernels' formulation (canonical orientation, class split, first applicable
    row) and the C restatement against the reference's own functions."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "rp_%s.npz" % name))
    t = z["table"]
    rows = G.rows_from_table(t)
    for what, (_, got) in (("program", rp_emul.cluster_interchr(lib, rows, z["chr_a"], z["chr_b"])),
                           ("oracle", rporacle.cluster_interchr(rows, z["chr_a"], z["chr_b"]))):
        for col, key in ((7, "pos_a"), (8, "pos_a1"), (9, "pos_b"), (10, "pos_b1"), (11, "visited"), (12, "report"), (13, "n_identical")):
            assert np.array_equal(got[key].astype(np.int64), t[:, col]), (what, key)
        tags, p = [], 0
        for c in z["tag_counts"]:
            tags.append(z["tag_values"][p:p + c].tolist())
            p += c
        holder = {x: i for i, tg in enumerate(tags) for x in tg}
        assert all(holder[int(t[i, 0])] == int(got["merged_into"][i]) for i in range(rows.size)), (what, "tags")
