"""GPU parity of the BGZF inflater (first half): pg_bgzf_inflate through the C ABI against
zlib -- the committed BAM fixture, BAM-like payloads at every compression level and strategy, and corrupt input."""
import gzip
import os
import zlib

import numpy as np
import pytest

from pindel_b200 import api, bgzf
from tests.test_bgzf_emulator import bam_like

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def pg():
    p = api.PatternGrowth()
    yield p
    p.close()


def test_committed_bam_fixture(pg):
    data = open(os.path.join(GOLD, "sim_prefix.bam"), "rb").read()
    before = pg.bgzf_launches
    got = pg.bgzf_inflate(data)
    assert pg.bgzf_launches == before + 1
    assert got.tobytes() == gzip.decompress(data)
    # the stream can also stay on the device
    assert pg.bgzf_inflate(data, to_host=False) == got.size
    ptr, n = pg.bgzf_device_data()
    assert ptr != 0 and n == got.size


@pytest.mark.parametrize("level", [0, 1, 6, 9])
@pytest.mark.parametrize("strategy", [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE])
def test_levels_and_strategies(pg, level, strategy):
    payload = bam_like(2_000_000, 17 + level)
    comp = bgzf.compress(payload, level=level, strategy=strategy, block=0xff00 if level else 0x8000)
    assert pg.bgzf_inflate(comp).tobytes() == payload


def test_payload_kinds(pg):
    rng = np.random.default_rng(3)
    for payload in (b"", b"A", b"A" * 500_000, bytes(rng.integers(0, 256, 300_000, dtype=np.uint8)),
                    bytes(rng.integers(0, 4, 300_000, dtype=np.uint8)), bytes(range(256)) * 1000):
        assert pg.bgzf_inflate(bgzf.compress(payload, block=0x7000)).tobytes() == payload
    p = 0.5 ** np.arange(1, 41)                     # codes longer than the fast tables
    payload = bytes(rng.choice(40, size=500_000, p=p / p.sum()).astype(np.uint8))
    assert pg.bgzf_inflate(bgzf.compress(payload, level=9)).tobytes() == payload


def test_corrupt_input_is_an_error(pg):
    payload = bam_like(300_000, 99)
    comp = bytearray(bgzf.compress(payload))
    with pytest.raises(api.PgError, match="not a BGZF"):
        pg.bgzf_inflate(b"this is not a bgzf file, not even close........")
    bad = bytearray(comp)
    bad[len(comp) // 2] ^= 0x10                      # inside some block's DEFLATE data
    with pytest.raises(api.PgError, match="corrupt"):
        pg.bgzf_inflate(bytes(bad))
    bad = bytearray(comp)
    k = bytes(comp).index(bgzf.EOF_MARKER) - 8       # CRC32 field of the last payload block
    bad[k] ^= 1
    with pytest.raises(api.PgError, match="CRC32"):
        pg.bgzf_inflate(bytes(bad))
    assert pg.bgzf_inflate(bytes(bad), check_crc=False).tobytes() == payload
    assert pg.bgzf_inflate(bytes(comp)).tobytes() == payload      # the context is still usable
