"""The BGZF inflater (pindel_b200/csrc/pg_bgzf.cuh, the per-thread program of pg_bgzf_inflate_kernel) compiled for the
host, against zlib: the committed BAM fixture, synthetic BAM-like payloads at every compression level and strategy
(stored, fixed and dynamic Huffman blocks, long codes that miss the fast tables), and corrupt input.  No GPU."""
import gzip
import os
import zlib

import numpy as np
import pytest

from pindel_b200 import bgzf
from tests.emul import bgzf_emul

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def lib():
    return bgzf_emul.build()


def bam_like(n, seed):
    from pindel_b200 import synth
    return synth.make_bam_like_bytes(n, seed)


def test_committed_bam_fixture(lib):
    data = open(os.path.join(GOLD, "sim_prefix.bam"), "rb").read()
    exp = gzip.decompress(data)                      # multi-member gzip = BGZF
    got, status = bgzf_emul.inflate(lib, data)
    assert not status.any(), status[status != 0][:5]
    assert got == exp and got[:4] == b"BAM\x01"


@pytest.mark.parametrize("level", [0, 1, 4, 6, 9])
@pytest.mark.parametrize("strategy", [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE])
def test_levels_and_strategies(lib, level, strategy):
    payload = bam_like(300_000, 7 + level)
    comp = bgzf.compress(payload, level=level, strategy=strategy, block=0xff00 if level else 0x8000)
    got, status = bgzf_emul.inflate(lib, comp)
    assert not status.any()
    assert got == payload


def test_payload_kinds(lib):
    rng = np.random.default_rng(3)
    for payload in (b"", b"A", b"A" * 200_000, bytes(rng.integers(0, 256, 150_000, dtype=np.uint8)),
                    bytes(rng.integers(0, 4, 150_000, dtype=np.uint8)), bytes(range(256)) * 300,
                    b"".join(bytes([i % 251]) * (i % 300) for i in range(1500))):
        # random bytes do not compress: small blocks keep BSIZE under 64 KiB
        got, status = bgzf_emul.inflate(lib, bgzf.compress(payload, block=0x7000))
        assert not status.any() and got == payload


def test_long_codes_miss_the_fast_tables(lib):
    """Skewed symbol frequencies give literal codes longer than 9 bits and distance codes longer than 6."""
    rng = np.random.default_rng(11)
    p = 0.5 ** np.arange(1, 41)
    syms = rng.choice(40, size=60_000, p=p / p.sum()).astype(np.uint8)
    payload = bytes(syms) + bytes(rng.integers(0, 256, 3000, dtype=np.uint8)) * 3
    c = zlib.compressobj(9, zlib.DEFLATED, -15, 9)
    raw = c.compress(payload[:65000]) + c.flush()
    rc, got = bgzf_emul.inflate_raw(lib, raw, 65000)
    assert rc == 0 and got == payload[:65000]


def test_corrupt_input_is_reported_not_crashed(lib):
    payload = bam_like(100_000, 99)
    comp = bytearray(bgzf.compress(payload))
    good, _ = bgzf_emul.inflate(lib, bytes(comp))
    assert good == payload
    rng = np.random.default_rng(5)
    flagged = 0
    for _ in range(200):
        bad = bytearray(comp)
        k = int(rng.integers(18, len(comp) - 40))
        bad[k] ^= 1 << int(rng.integers(0, 8))
        try:
            got, status = bgzf_emul.inflate(lib, bytes(bad))
        except ValueError:
            flagged += 1                  # a header byte: the walk refuses the file
            continue
        if status.any() or got != payload:
            flagged += 1
        if not status.any():
            # whatever the inflater accepted (CRC included) must be what zlib makes of the same bytes
            assert gzip.decompress(bytes(bad)) == got
    assert flagged > 190                  # (a flip in MTIME/XFL/OS header bytes changes nothing)
    with pytest.raises(ValueError):
        bgzf_emul.inflate(lib, b"not a bgzf file at all........")


class _Bits:
    """LSB-first bit writer (RFC 1951 3.1.1); Huffman codes go in MSB first."""
    def __init__(self):
        self.acc, self.n, self.out = 0, 0, bytearray()

    def put(self, v, nbits):
        self.acc |= v << self.n
        self.n += nbits
        while self.n >= 8:
            self.out.append(self.acc & 255)
            self.acc >>= 8
            self.n -= 8

    def code(self, v, nbits):
        for i in range(nbits - 1, -1, -1):
            self.put((v >> i) & 1, 1)

    def bytes(self):
        return bytes(self.out) + (bytes([self.acc & 255]) if self.n else b"")


def _canonical(lengths):
    codes, code = {}, 0
    for l in range(1, 16):
        for s, sl in enumerate(lengths):
            if sl == l:
                codes[s] = (code, l)
                code += 1
        code <<= 1
    return codes


def _dynamic_block(lit_len, dist_len, symbols, cl_len=None):
    """One final dynamic block whose code sets are given as they are (complete or not).  The lengths are sent plainly
    (code-length symbols 0..15 only).  cl_len: the 19 lengths of the code-length code, default: 4 bits for 0..15."""
    b = _Bits()
    b.put(1, 1)
    b.put(2, 2)
    nlen, ndist = max(257, len(lit_len)), max(1, len(dist_len))
    lit_len = list(lit_len) + [0] * (nlen - len(lit_len))
    dist_len = list(dist_len) + [0] * (ndist - len(dist_len))
    if cl_len is None:
        cl_len = [4] * 16 + [0, 0, 0]
    order = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
    b.put(nlen - 257, 5)
    b.put(ndist - 1, 5)
    b.put(19 - 4, 4)
    for s in order:
        b.put(cl_len[s], 3)
    clc = _canonical(cl_len)
    for l in lit_len + dist_len:
        b.code(*clc[l])
    lc, dc = _canonical(lit_len), _canonical(dist_len)
    for kind, s, extra, nextra in symbols:
        b.code(*(lc if kind == "L" else dc)[s])
        if nextra:
            b.put(extra, nextra)
    b.put(0, 16)
    return b.bytes()


def _zlib_raw(raw):
    try:
        return zlib.decompressobj(-15).decompress(raw)
    except zlib.error:
        return None


def test_incomplete_code_sets_are_judged_like_zlib(lib):
    """zlib's inflate_table refuses an incomplete set of code lengths, except a single length-1 code of the
    literal/length or distance code (and the empty set); pgz_build follows it, so a damaged dynamic block is rejected
    here exactly when zlib rejects it, with or without the CRC."""
    A, B, EOB = 65, 66, 256
    lit = [0] * 257

    def lits(**kw):
        v = list(lit)
        for s, l in kw.items():
            v[int(s[1:])] = l
        return v
    text = [("L", A, 0, 0), ("L", B, 0, 0), ("L", A, 0, 0), ("L", EOB, 0, 0)]
    cases = {
        # complete literal code {A:1, B:2, EOB:2}, single distance code of length 1: accepted by both
        "complete+single_dist": (lits(s65=1, s66=2, s256=2), [1], text),
        # ... and the distance code is used (length 3 = code 257 would need a length code; use a longer complete set)
        "complete+no_dist": (lits(s65=1, s66=2, s256=2), [0], text),
        # incomplete literal code {A:2, B:2, EOB:2}: zlib says "invalid literal/lengths set"
        "incomplete_lit": (lits(s65=2, s66=2, s256=2), [1], text),
        # incomplete distance code {0:2, 1:2}: "invalid distances set"
        "incomplete_dist": (lits(s65=1, s66=2, s256=2), [2, 2], text),
        # single distance code of length 2: incomplete and not the exception
        "single_dist_len2": (lits(s65=1, s66=2, s256=2), [2], text),
        # the only literal/length code is the end of block, one bit: zlib's exception applies to this tree too
        "single_lit_eob": (lits(s256=1), [1], [("L", EOB, 0, 0)]),
        # over-subscribed literal code
        "oversubscribed_lit": (lits(s65=1, s66=1, s256=1), [1], [("L", EOB, 0, 0)]),
    }
    seen = set()
    for name, (ll, dl, syms) in cases.items():
        try:
            raw = _dynamic_block(ll, dl, syms)
        except KeyError:
            raise AssertionError(name)
        exp = _zlib_raw(raw)
        rc, got = bgzf_emul.inflate_raw(lib, raw, 16 if exp is None else len(exp))   # (ISIZE is part of a BGZF block)
        assert (rc == 0) == (exp is not None), (name, rc, exp)
        if exp is not None:
            assert got == exp, name
        else:
            assert rc == 1, (name, rc)            # PGZ_ERR_HEADER: refused when the tables are built, not further on
        seen.add(exp is not None)
    assert seen == {True, False}
    # the code-length code itself: an incomplete one (sixteen 5-bit codes) is refused even though every length decodes
    raw = _dynamic_block(lits(s65=1, s66=2, s256=2), [1], text, cl_len=[5] * 16 + [0, 0, 0])
    assert _zlib_raw(raw) is None
    assert bgzf_emul.inflate_raw(lib, raw, 3)[0] == 1
    # a match through the single distance code: both decoders produce the same bytes
    ll = lits(s65=2, s256=2)
    ll += [2, 2]                                  # 257 (length 3), 258 (length 4)
    raw = _dynamic_block(ll, [1], [("L", A, 0, 0), ("L", 257, 0, 0), ("D", 0, 0, 0), ("L", EOB, 0, 0)])
    exp = _zlib_raw(raw)
    rc, got = bgzf_emul.inflate_raw(lib, raw, 4)
    assert exp == b"AAAA" and rc == 0 and got == exp


def test_damaged_streams_are_judged_like_zlib_without_the_crc(lib):
    """scripts/bgzf_corruption_campaign.py in small: bit flips, overwritten bytes, truncations and random block headers in raw
    DEFLATE streams of every block type -- zlib inflates the damaged stream completely to ISIZE bytes <=> the inflater returns
    PGZ_OK with the same bytes; nothing is decoded from beyond the end of the input (a stored block whose LEN reaches past a
    cut-off stream was once filled with the zeros the bit buffer delivers there)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import bgzf_corruption_campaign as camp
    rng = np.random.default_rng(77)
    streams = []
    for level, strategy in ((1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (0, zlib.Z_DEFAULT_STRATEGY)):
        for size in (300, 5000):
            payload = bam_like(size, 300 + level + size)
            c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
            streams.append((c.compress(payload) + c.flush(), payload))
    both_ok = both_bad = 0
    for _ in range(6000):
        raw, payload = streams[int(rng.integers(0, len(streams)))]
        bad = camp.damage(rng, raw)
        exp = camp.zlib_raw(bad, len(payload))
        rc, got = bgzf_emul.inflate_raw(lib, bad, len(payload))
        assert (rc == 0) == (exp is not None) and (exp is None or got == exp), (rc, exp is None)
        both_ok += exp is not None
        both_bad += exp is None
    assert both_ok > 500 and both_bad > 2000
    # the stored block cut off in the middle: LEN says 5000, 100 bytes are there
    raw, payload = streams[-1]
    assert raw[0] & 6 == 0 and bgzf_emul.inflate_raw(lib, raw, len(payload)) == (0, payload)
    rc, _ = bgzf_emul.inflate_raw(lib, raw[:105], len(payload))
    assert rc == 5 and camp.zlib_raw(raw[:105], len(payload)) is None          # PGZ_ERR_INPUT
