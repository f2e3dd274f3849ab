"""The reader's split-read filter as the device runs it (pindel_b200/csrc/pg_sr.cuh compiled for the host, tests/emul/sr_emul.cpp:
analysis, name table, partner by rank, prefix counts, scatter) against the reference's own dumps (tests/golden/sr_*.npz) and the
Python oracle.  The GPU run of the same functions is pending."""
import gzip
import os

import numpy as np
import pytest

from oracle import sroracle
from pindel_b200 import bam
from tests.emul import bam_emul, sr_emul
from tests.golden import make_sr_golden as G

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def case_stream(case):
    if case == "sim_prefix":
        return gzip.decompress(open(os.path.join(GOLD, "sim_prefix.bam"), "rb").read())
    return G.paired_stream(**G.SYNTH)


def rows_of_product(stream, recs, cands, seqs, refs, insert_size):
    """product output -> the layout of the reference dump / oracle.sroracle.candidates"""
    out = []
    for c, s in zip(cands, seqs):
        u = sroracle.Rec(stream, int(recs["off"][c["read"]]))
        name = b"@" + u.qname + (b"/1" if u.flag & 0x40 else b"/2" if u.flag & 0x80 else b"")
        out.append((sroracle.fnv1a(name), int(c["matched_d"]), int(c["matched_rel_pos"]), int(c["ms"]), insert_size, int(c["read_length"]),
                    sroracle.fnv1a(s[:c["read_length"]].tobytes())))
    r = np.stack([refs["pos"].astype(np.int64), refs["mq"].astype(np.int64), refs["read_length"].astype(np.int64)], axis=1)
    return np.array(out, dtype=np.int64).reshape(-1, 7), r.reshape(-1, 3)


@pytest.fixture(scope="module")
def libs():
    return sr_emul.build(), bam_emul.build()


@pytest.mark.parametrize("case", list(G.QUERIES))
def test_device_functions_match_reference_dumps(libs, case):
    lib, bl = libs
    z = np.load(os.path.join(GOLD, "sr_%s.npz" % case))
    stream = case_stream(case)
    recs, _ = bam_emul.index(bl, stream)
    for k, (tid, beg, end, isz, min_q) in enumerate(G.QUERIES[case]):
        q = sr_emul.query(tid, beg, end, isz, int(z["chr_len"][tid]), min_q)
        # insertion order of the name table (arbitrary on the device), whole-key collisions (8 bits of hash left)
        for seed, hash_mask in ((0, 0), (5, 0), (7, 0xFF << 20)):
            cands, seqs, refs, _, err = sr_emul.candidates(lib, stream, recs, q, seed, 0, hash_mask)
            assert err == 0
            c, r = rows_of_product(stream, recs, cands, seqs, refs, isz)
            assert np.array_equal(c, z["cand_%d" % k]), ("candidates", case, k, seed, hash_mask)
            assert np.array_equal(r, z["refs_%d" % k]), ("reference-supporting reads", case, k, seed, hash_mask)


def named_stream(names, seed=3, weird=True):
    """records at increasing positions with the given names, all clipped (weird) or all plain"""
    rng = np.random.default_rng(seed)
    out = [bam.header([("chr1", 100_000)])]
    for i, nm in enumerate(names):
        seq = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 60))
        cig = [(4, 20), (0, 40)] if weird else [(0, 60)]
        flag = 1 | (0x40 if i % 2 == 0 else 0x80) | (16 if i % 3 == 0 else 0)
        body = bam.record(0, 1000 + 10 * i, nm, flag, 40, cig, 60, 0, 1000, 0, b"NMC\x01", rng)
        # bam.record draws random sequence nibbles: overwrite them with plain bases
        l_name = len(nm) + 1
        off = 4 + 32 + l_name + 4 * len(cig)
        body = body[:off] + G.pack_seq(seq) + body[off + 30:]
        out.append(body)
    return b"".join(out)


@pytest.mark.parametrize("names", [
    [b"same"] * 7,                                           # ranks 0..6: pairs (0,1) (2,3) (4,5), the last one waits
    [b"a", b"ab", b"a", b"abc", b"ab", b"abc", b"a"],        # names that are prefixes of one another
    [b"r%d" % (i % 5) for i in range(23)],
], ids=["one-name", "prefixes", "five-names"])
def test_pairing_by_rank_matches_the_streaming_map(libs, names):
    lib, bl = libs
    stream = named_stream(names)
    table = bam.serial_index(stream)
    recs, _ = bam_emul.index(bl, stream)
    exp_c, exp_r = sroracle.candidates(stream, table, 0, 0, 100_000, 500, 100_000)
    assert exp_c.shape[0] > len(names)
    for seed, bits, hash_mask in ((0, 0, 0), (3, 5, 0), (9, 6, 1 << 20), (11, 0, 0xFFFF)):
        cands, seqs, refs, delivered, err = sr_emul.candidates(lib, stream, recs, sr_emul.query(0, 0, 100_000, 500, 100_000), seed, bits, hash_mask)
        c, r = rows_of_product(stream, recs, cands, seqs, refs, 500)
        assert delivered == len(names) and err == 0
        assert np.array_equal(c, exp_c) and np.array_equal(r, exp_r), (seed, bits, hash_mask)


def test_insert_size_not_above_the_read_length_is_reported(libs):
    """the reference stops the program there (src/reader.cpp:866-874); the device reports it and writes nothing"""
    lib, bl = libs
    stream = named_stream([b"x", b"y", b"x"])
    recs, _ = bam_emul.index(bl, stream)
    _, _, _, _, err = sr_emul.candidates(lib, stream, recs, sr_emul.query(0, 0, 100_000, 60, 100_000))
    assert err != 0
    _, _, _, _, err = sr_emul.candidates(lib, stream, recs, sr_emul.query(0, 0, 100_000, 61, 100_000))
    assert err == 0
    with pytest.raises(ValueError):
        sroracle.candidates(stream, bam.serial_index(stream), 0, 0, 100_000, 60, 100_000)


def test_plain_pairs_give_reference_reads_only(libs):
    lib, bl = libs
    names = [b"p%d" % (i // 2) for i in range(20)]
    stream = named_stream(names, weird=False)
    stream = stream.replace(b"NMC\x01", b"NMC\x00")
    recs, _ = bam_emul.index(bl, stream)
    cands, seqs, refs, delivered, err = sr_emul.candidates(lib, stream, recs, sr_emul.query(0, 0, 100_000, 500, 100_000))
    exp_c, exp_r = sroracle.candidates(stream, bam.serial_index(stream), 0, 0, 100_000, 500, 100_000)
    assert cands.size == 0 and exp_c.shape[0] == 0 and refs.size == 20 and exp_r.shape[0] == 20
    assert np.array_equal(rows_of_product(stream, recs, cands, seqs, refs, 500)[1], exp_r)


def test_malformed_record_in_the_window_is_an_error(libs):
    """a record whose CIGAR count points past its block_size: never looked into, reported; outside the window it is harmless"""
    import struct
    lib, bl = libs
    stream = bytearray(named_stream([b"x", b"y", b"x", b"z"]))
    recs, _ = bam_emul.index(bl, bytes(stream))
    off = int(recs["off"][2])
    struct.pack_into("<H", stream, off + 4 + 12, 60000)            # n_cigar_op
    stream = bytes(stream)
    recs, _ = bam_emul.index(bl, stream)
    assert recs.size == 4
    _, _, _, _, err = sr_emul.candidates(lib, stream, recs, sr_emul.query(0, 0, 100_000, 500, 100_000), 3)
    assert err & 128
    cands, _, _, delivered, err = sr_emul.candidates(lib, stream, recs, sr_emul.query(0, 0, 1015, 500, 100_000))       # records 0 and 1 only
    assert err == 0 and delivered == 2 and cands.size == 2
