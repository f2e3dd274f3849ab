"""scripts/bgzf_corruption_campaign.py -- CPU-only: the inflater of pg_bgzf.cuh (host build, tests/emul/bgzf_emul.cpp) against zlib
on DAMAGED raw DEFLATE streams, without the CRC: random bit flips, byte overwrites, truncations and random Huffman headers in the
payloads of BGZF blocks (all block types, levels 1-9, the four strategies).  For every damaged stream both decoders must agree:
zlib inflates it completely to exactly ISIZE bytes <=> the inflater returns PGZ_OK, and then the bytes are equal; otherwise the
inflater must return an error (never crash, never write past ISIZE).  This is the behaviour htslib's inflate_block has on a damaged
block before it looks at the CRC.      python scripts/bgzf_corruption_campaign.py [--rounds N] [--seed S]"""
import argparse
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def zlib_raw(raw, isize):
    d = zlib.decompressobj(-15)
    try:
        out = d.decompress(raw, isize + 1)
    except zlib.error:
        return None
    if not d.eof or len(out) != isize:
        return None
    return out


def damage(rng, raw):
    b = bytearray(raw)
    kind = int(rng.integers(0, 5))
    if kind == 0:                                   # one bit
        k = int(rng.integers(0, len(b)))
        b[k] ^= 1 << int(rng.integers(0, 8))
    elif kind == 1:                                 # a few bits near the block header (code-length code, code lengths)
        for _ in range(int(rng.integers(1, 4))):
            k = int(rng.integers(0, min(len(b), 80)))
            b[k] ^= 1 << int(rng.integers(0, 8))
    elif kind == 2:                                 # a byte overwritten
        b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
    elif kind == 3:                                 # truncated
        b = b[:int(rng.integers(1, len(b)))]
    else:                                           # a random dynamic-block header in front of the rest
        n = int(rng.integers(4, 40))
        b[:n] = bytes([0x05 | (int(rng.integers(0, 32)) << 3)]) + bytes(rng.integers(0, 256, n - 1, dtype=np.uint8))
    return bytes(b)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=20000)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    from pindel_b200 import synth
    from tests.emul import bgzf_emul
    lib = bgzf_emul.build()
    rng = np.random.default_rng(a.seed)
    streams = []
    for level, strategy in ((1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED),
                            (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE), (0, zlib.Z_DEFAULT_STRATEGY)):
        for size in (300, 4000, 30000):
            payload = synth.make_bam_like_bytes(size, 100 + level + size)
            c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
            streams.append((c.compress(payload) + c.flush(), payload))
    accepted = rejected = 0
    for r in range(a.rounds):
        raw, payload = streams[int(rng.integers(0, len(streams)))]
        bad = damage(rng, raw)
        exp = zlib_raw(bad, len(payload))
        rc, got = bgzf_emul.inflate_raw(lib, bad, len(payload))
        if exp is None:
            assert rc != 0, "round %d: zlib refuses the stream, the inflater accepted it (%d bytes)" % (r, len(got))
            rejected += 1
        else:
            assert rc == 0 and got == exp, "round %d: zlib inflates the stream, the inflater says %d" % (r, rc)
            accepted += 1
    for raw, payload in streams:                    # and the undamaged ones
        assert bgzf_emul.inflate_raw(lib, raw, len(payload)) == (0, payload)
    print("campaign clean: %d damaged streams, %d refused by both decoders, %d inflated by both to the same bytes" % (a.rounds, rejected, accepted))


if __name__ == "__main__":
    main()
