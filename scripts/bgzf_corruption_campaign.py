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


def files(a):
    """Whatever the walk and the inflater accept with the CRC checked must be what gzip makes of the same bytes."""
    import gzip
    from pindel_b200 import bgzf, synth
    from tests.emul import bgzf_emul
    lib = bgzf_emul.build()
    rng = np.random.default_rng(a.seed)
    originals = [bgzf.compress(synth.make_bam_like_bytes(n, 7 + n), level=lv, block=blk) for n, lv, blk in
                 ((200, 6, 0xff00), (70_000, 6, 0x4000), (150_000, 1, 0xff00), (40_000, 0, 0x3000))]
    refused = flagged = same = 0
    for r in range(a.rounds):
        b = bytearray(originals[int(rng.integers(0, len(originals)))])
        for _ in range(int(rng.integers(1, 4))):
            kind = int(rng.integers(0, 4))
            if kind == 0:                               # somewhere in a block header: the first 18 bytes after a block start
                starts, p = [], 0
                while p + 18 <= len(b) and b[p] == 31:
                    starts.append(p)
                    p += (b[p + 16] | b[p + 17] << 8) + 1
                p = starts[int(rng.integers(0, len(starts)))] if starts else 0
                if p + 18 <= len(b):
                    b[p + int(rng.integers(0, 18))] = int(rng.integers(0, 256))
            elif kind == 1 and len(b) > 1:
                b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
            elif kind == 2 and len(b) > 30:
                b = b[:int(rng.integers(0, len(b)))]
            elif len(b) > 8:                            # the trailer (CRC32, ISIZE) of the last block
                b[len(b) - 1 - int(rng.integers(0, 8))] = int(rng.integers(0, 256))
        bad = bytes(b)
        try:
            got, status = bgzf_emul.inflate(lib, bad)
        except ValueError:
            refused += 1
            continue
        if status.any():
            flagged += 1
            continue
        # htslib's check_header looks at FLG & 4 only (bgzf.c), gzip honours the other flag bits (FNAME, FCOMMENT, FHCRC ...): a file
        # accepted here is compared with what gzip makes of it once every block's FLG is what BGZF writes
        norm, p = bytearray(bad), 0
        while p + 18 <= len(norm) and norm[p] == 31:
            norm[p + 3] = 4
            p += (norm[p + 16] | norm[p + 17] << 8) + 1 if (norm[p + 10] | norm[p + 11] << 8) == 6 else len(norm)
        try:
            exp = gzip.decompress(bytes(norm))
        except Exception as e:
            raise AssertionError("round %d: accepted a file gzip refuses (%s)" % (r, e))
        assert got == exp, "round %d: accepted bytes differ from gzip's" % r
        same += 1
    print("campaign clean: %d damaged BGZF files: %d refused by the header walk, %d by a block's status, %d accepted and equal to gzip's output" %
          (a.rounds, refused, flagged, same))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=20000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--files", action="store_true", help="whole BGZF files damaged anywhere, block headers (BSIZE, XLEN, the extra "
                    "subfields, ISIZE, CRC) included, and cut off at any byte: the host's header walk (pg_bgzf_host.h) and the inflater")
    a = ap.parse_args()
    if a.files:
        return files(a)
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
