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
        assert lib.bamemu_last_walks() == stats[0] + stats[1]      # a repair walks its segment alone (pgb_settle)
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


def run_expectation(table, stream_len, lo, hi):
    """what the index of the run stream[lo:hi] must be: the records that start at or after lo and fit before hi, offsets
    relative to lo; the offset of the first one that does not fit (or the run's length)"""
    size = np.append(table["off"][1:], stream_len) - table["off"]
    inside = (table["off"] >= lo) & (table["off"] + size <= hi)
    first = int(table["off"][table["off"] >= lo][0]) - lo if (table["off"] >= lo).any() else hi - lo
    exp = table[inside].copy()
    exp["off"] -= lo
    after = table["off"][(table["off"] >= lo) & ~inside]
    tail = int(after[0]) - lo if after.size else hi - lo
    return exp, first, min(tail, hi - lo)


@pytest.mark.parametrize("adversarial", [False, True], ids=["plain", "fake-records-in-tags"])
def test_runs_of_blocks(lib, adversarial):
    """a window's chunk instead of the whole file: no header in front, the last record cut off by the end of the run"""
    stream = bam.synthetic_stream(3000, 8, adversarial=adversarial)
    table = bam.serial_index(stream)
    rng = np.random.default_rng(4)
    block = 0x4000
    n_blocks = -(-len(stream) // block)
    cut, redone = 0, 0
    for _ in range(40):
        b0 = int(rng.integers(1, n_blocks - 1))
        b1 = int(rng.integers(b0 + 1, n_blocks + 1))
        lo, hi = b0 * block, min(b1 * block, len(stream))
        exp, first, tail = run_expectation(table, len(stream), lo, hi)
        for seg in (900, 5000, 0):
            got, got_tail, again = bam_emul.index_run(lib, stream[lo:hi], first, 2, seg)
            assert same(got, exp) and got_tail == tail, (b0, b1, seg, got.size, exp.size, got_tail, tail)
            redone += again
        cut += tail < hi - lo
    assert cut > 20                       # most runs end inside a record
    if adversarial:
        assert redone > 0
    # a run that ends on a record boundary, a run that is one partial record, an impossible block_size
    k = table.size // 2
    lo, hi = int(table["off"][k]), int(table["off"][k + 50])
    got, tail, _ = bam_emul.index_run(lib, stream[lo:hi], 0, 2, 700)
    assert got.size == 50 and tail == hi - lo
    got, tail, _ = bam_emul.index_run(lib, stream[lo:lo + 20], 0, 2)
    assert got.size == 0 and tail == 0
    k = table.size // 4                                          # refID 0 here: the four bytes after block_size are zero
    lo, hi = int(table["off"][k]), int(table["off"][k + 50])
    with pytest.raises(ValueError):
        bam_emul.index_run(lib, stream[lo + 4:hi], 0, 2)          # first_record is not a record start: block_size 0


def test_window_queries_on_a_run_equal_those_on_the_whole_file(lib):
    """discovery and the split-read filter on the table of a run give what they give on the whole file, when the window's
    records lie inside the run"""
    from tests.emul import sr_emul
    from tests.golden import make_sr_golden as G
    srl = sr_emul.build()
    stream = G.paired_stream(4000, 9, refs=G.SYNTH_REFS[:1])
    table = bam.serial_index(stream)
    whole, _ = bam_emul.index(lib, stream)
    block = 0x8000
    lo, hi = 6 * block, 20 * block
    exp, first, tail = run_expectation(table, len(stream), lo, hi)
    run = stream[lo:hi]
    recs, got_tail, _ = bam_emul.index_run(lib, run, first, 1)
    assert same(recs, exp) and got_tail == tail
    # a window whose delivered records (and the one that ends the iteration) are all inside the run
    inside = table[(table["off"] >= lo + first) & (table["off"] < lo + tail)]
    beg, end = int(inside["pos"][50]), int(inside["pos"][-50])
    a = bam_emul.rp_discovery(lib, stream, whole, 0, beg, end, 500)
    b = bam_emul.rp_discovery(lib, run, recs, 0, beg, end, 500)
    assert a[4] == b[4] and all(np.array_equal(x, y) for x, y in ((a[0], b[0]), (a[2], b[2])))
    q = sr_emul.query(0, beg, end, 500, 400_000)
    ca, sa, ra, da, _ = sr_emul.candidates(srl, stream, whole, q)
    cb, sb, rb, db, _ = sr_emul.candidates(srl, run, recs, q)
    assert da == db and ca.size == cb.size > 100 and np.array_equal(sa, sb) and np.array_equal(ra, rb)
    shift = int(np.nonzero(whole["off"] == lo + first)[0][0])          # record indices differ by the records in front of the run
    for k in ("matched_rel_pos", "read_length", "matched_d", "ms"):
        assert np.array_equal(ca[k], cb[k]), k
    assert np.array_equal(ca["anchor"], cb["anchor"] + shift) and np.array_equal(ca["read"], cb["read"] + shift)
