"""GPU parity of the discordant-pair discovery (first half): pg_bgzf_inflate + pg_bam_index +
pg_bam_rp_discovery through the C ABI against the reference's own dumps (tests/golden/rpd_*.npz) and the oracle."""
import os

import numpy as np
import pytest

from oracle import rpdoracle
from pindel_b200 import api, bam, bgzf
from tests.golden import make_rpd_golden as G
from tests.test_rp_discovery import case_stream, rows_of_product

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def pg():
    p = api.PatternGrowth()
    yield p
    p.close()


@pytest.mark.parametrize("case", list(G.QUERIES))
def test_matches_reference_dumps(pg, case):
    z = np.load(os.path.join(GOLD, "rpd_%s.npz" % case))
    stream = case_stream(case)
    data = open(os.path.join(GOLD, "sim_prefix.bam"), "rb").read() if case == "sim_prefix" else bgzf.compress(stream)
    pg.bgzf_inflate(data, to_host=False)
    recs, _, _ = pg.bam_index()
    for k, (tid, beg, end, isz) in enumerate(G.QUERIES[case]):
        rows, src, irows, isrc, scanned = pg.bam_rp_discovery(tid, beg, end, isz)
        assert scanned == int(z["scanned_%d" % k])
        assert np.array_equal(rows_of_product(stream, recs, rows, src, irows, isrc), z["rows_%d" % k]), (case, k)


def test_matches_oracle_on_a_fresh_stream_and_quality_threshold(pg):
    stream = bam.synthetic_stream(40_000, 91, discordant=0.3)
    table = bam.serial_index(stream)
    pg.bgzf_inflate(bgzf.compress(stream), to_host=False)
    recs, _, _ = pg.bam_index()
    for q in ((0, 0, 50_000_000, 500, 0), (0, 100_000, 400_000, 250, 30), (1, 0, 30_000_000, 500, 20), (0, 5_000, 5_100, 500, 0)):
        exp, _, sc = rpdoracle.discovery(stream, table, *q)
        rows, src, irows, isrc, scanned = pg.bam_rp_discovery(*q)
        assert scanned == sc and np.array_equal(rows_of_product(stream, recs, rows, src, irows, isrc), exp), q


def test_needs_an_indexed_stream(pg):
    fresh = api.PatternGrowth()
    with pytest.raises(api.PgError):
        fresh.bam_rp_discovery(0, 0, 1000, 500)
    fresh.bgzf_inflate(bgzf.compress(bam.synthetic_stream(100, 3)), to_host=False)
    with pytest.raises(api.PgError):
        fresh.bam_rp_discovery(0, 0, 1000, 500)          # inflated but not indexed
    fresh.close()
