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
