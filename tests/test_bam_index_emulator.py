"""The BAM record index (pindel_b200/csrc/pg_bam.cuh + pg_bam_host.h: segment guesses, walks, consistency check) compiled
for the host, against the plain serial walk: the committed BAM fixture, synthetic streams at several segment sizes, and
streams whose tag data contains whole fake records (the guess must be caught by the check and walked again).  No GPU."""
import gzip
import os

import numpy as np
import pytest

from pindel_b200 import bam
from tests.emul import bam_emul

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def lib():
    return bam_emul.build()


def same(a, b):
    return a.size == b.size and all(np.array_equal(a[k], b[k]) for k in ("off", "tid", "pos", "end", "flag", "mapq"))


def test_committed_bam_fixture(lib):
    stream = gzip.decompress(open(os.path.join(GOLD, "sim_prefix.bam"), "rb").read())
    exp = bam.serial_index(stream)
    for seg in (0, 4096, 50_000):
        got, stats = bam_emul.index(lib, stream, seg)
        assert same(got, exp), seg
        assert stats[1] == 0          # real data: every guess was right
    assert exp.size > 10_000


@pytest.mark.parametrize("seg", [512, 3000, 65536, 0])
def test_synthetic_streams(lib, seg):
    stream = bam.synthetic_stream(4000, 5)
    got, stats = bam_emul.index(lib, stream, seg)
    assert same(got, bam.serial_index(stream))
    assert stats[0] == max(1, -(-(len(stream) - stats[2]) // (seg or 256 << 10)))


def test_fake_records_inside_tags_are_caught(lib):
    stream = bam.synthetic_stream(3000, 6, adversarial=True)
    exp = bam.serial_index(stream)
    redone = 0
    for seg in (700, 1500, 4000):
        got, stats = bam_emul.index(lib, stream, seg)
        assert same(got, exp), seg
        redone += stats[1]
    assert redone > 0                  # some segment boundary fell into a tag and locked on to the fake records


def test_degenerate_streams(lib):
    hdr = bam.header([("c", 1000)])
    got, _ = bam_emul.index(lib, hdr)                 # header only
    assert got.size == 0
    one = hdr + bam.record(0, 5, b"r", 0, 30, [(0, 10)], 10)
    got, _ = bam_emul.index(lib, one, 16)
    assert same(got, bam.serial_index(one)) and got.size == 1
    with pytest.raises(ValueError):
        bam_emul.index(lib, b"not a bam stream....")
    with pytest.raises(ValueError):
        bam_emul.index(lib, one[:-3])                 # truncated record
