"""The reader's split-read filter (rest of SURVEY.md section 8(f) row N3): oracle/sroracle.py against the reference's own
ReadInBamReads_SR (src/reader.cpp:483-560) -- its committed dumps (tests/golden/sr_*.npz, made through oracle/ref_hook.cpp
pgref_sr_candidates by tests/golden/make_sr_golden.py).  Oracle-first groundwork: no device stage consumes it yet."""
import gzip
import os

import numpy as np
import pytest

from oracle import sroracle
from pindel_b200 import bam
from tests.golden import make_sr_golden as G

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def case_stream(case):
    if case == "sim_prefix":
        return gzip.decompress(open(os.path.join(GOLD, "sim_prefix.bam"), "rb").read())
    return G.paired_stream(**G.SYNTH)


@pytest.mark.parametrize("case", list(G.QUERIES))
def test_oracle_matches_reference_dumps(case):
    z = np.load(os.path.join(GOLD, "sr_%s.npz" % case))
    stream = case_stream(case)
    table = bam.serial_index(stream)
    nonempty = 0
    for k, (tid, beg, end, isz, min_q) in enumerate(G.QUERIES[case]):
        cand, refs = sroracle.candidates(stream, table, tid, beg, end, isz, int(z["chr_len"][tid]), min_anchor_quality=min_q)
        assert np.array_equal(cand, z["cand_%d" % k]), ("candidates", case, k)
        assert np.array_equal(refs, z["refs_%d" % k]), ("reference-supporting reads", case, k)
        nonempty += cand.shape[0] > 0
    assert nonempty >= 2


def test_synthetic_stream_exercises_the_branches():
    """the seeded stream reaches every outcome of the filter: self-anchored clipped reads, unmapped mates behind an anchor,
    reads dropped for their N content or length, reverse-strand anchors, anchors refused for their flags"""
    stream = G.paired_stream(**G.SYNTH)
    table = bam.serial_index(stream)
    recs = [sroracle.Rec(stream, int(o)) for o in table["off"]]
    fl = [sroracle.flags_of(r) for r in recs]
    weird = [sroracle.is_weird_read(f, r) for f, r in zip(fl, recs)]
    assert 0 < sum(weird) < len(recs)
    assert any(not f["mapped"] for f in fl)
    assert any(sroracle.build_record_sr(r, r, 500, 400_000, 0) is None for r, w in zip(recs, weird) if w)
    assert any(not sroracle.is_good_anchor(f, r, 30) for f, r in zip(fl, recs) if f["mapped"] and r.qual >= 30)
    assert any(sroracle.is_ref_read(f, r, 0.02, 2) for f, r in zip(fl, recs))
    assert any(r.flag & 16 and w for r, w in zip(recs, weird))
