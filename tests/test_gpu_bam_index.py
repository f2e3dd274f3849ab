"""GPU parity of the BAM record index: pg_bgzf_inflate + pg_bam_index through the C ABI
against the plain serial walk -- the committed BAM fixture, synthetic streams spanning many segments, streams whose tag
data contains whole fake records, and broken input."""
import gzip
import os

import numpy as np
import pytest

from pindel_b200 import api, bam, bgzf
from tests.test_bam_index_emulator import same

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def pg():
    p = api.PatternGrowth()
    yield p
    p.close()


def test_committed_bam_fixture(pg):
    data = open(os.path.join(GOLD, "sim_prefix.bam"), "rb").read()
    stream = gzip.decompress(data)
    pg.bgzf_inflate(data, to_host=False)
    before = pg.bam_launches
    recs, first, n_ref = pg.bam_index()
    assert pg.bam_launches > before
    exp = bam.serial_index(stream)
    assert same(recs, exp) and (first, n_ref) == bam.first_record_offset(stream) and first == int(exp["off"][0])
    assert pg.bam_last_segments()[0] == -(-(len(stream) - first) // (256 << 10))


@pytest.mark.parametrize("n_records,seed,adversarial", [(30_000, 7, False), (20_000, 8, True)])
def test_synthetic_streams(pg, n_records, seed, adversarial):
    stream = bam.synthetic_stream(n_records, seed, adversarial=adversarial)
    pg.bgzf_inflate(bgzf.compress(stream), to_host=False)
    recs, _, _ = pg.bam_index()
    assert same(recs, bam.serial_index(stream))
    assert pg.bam_last_segments()[0] > 10


def test_errors(pg):
    fresh = api.PatternGrowth()
    with pytest.raises(api.PgError):
        fresh.bam_index()                                    # nothing inflated
    fresh.close()
    pg.bgzf_inflate(bgzf.compress(b"this is not a BAM stream, just text " * 100), to_host=False)
    with pytest.raises(api.PgError, match="not a BAM"):
        pg.bam_index()
    stream = bam.synthetic_stream(500, 9)
    pg.bgzf_inflate(bgzf.compress(stream[:-5]), to_host=False)
    with pytest.raises(api.PgError, match="truncated"):
        pg.bam_index()
    pg.bgzf_inflate(bgzf.compress(bam.header([("c", 100)])), to_host=False)
    recs, first, n_ref = pg.bam_index()
    assert recs.size == 0 and n_ref == 1
