"""Random BAM streams for the split-read filter (tests/test_sr_fuzz.py, scripts/cpu_fuzz_campaign.py): named mate pairs with the
format's corners -- every CIGAR operation, NM tags of every integer type (and a string-typed one), other tags of every type in
front of NM, no tags at all, empty sequences, mapped reads without a CIGAR, long names with common prefixes, names occurring
one to five times, all flag bits."""
from __future__ import annotations

import struct

import numpy as np

from pindel_b200 import bam
from tests.golden.make_sr_golden import pack_seq

REFS = (("chr1", 200_000), ("chr2", 150_000))


def _aux(rng) -> bytes:
    out = b""
    for _ in range(int(rng.integers(0, 4))):           # tags the walker must step over
        k = int(rng.integers(0, 8))
        tag = bytes(rng.choice(np.frombuffer(b"XYZ", dtype=np.uint8), 1)) + bytes(rng.choice(np.frombuffer(b"012abc", dtype=np.uint8), 1))
        if k == 0:
            out += tag + b"A" + b"U"
        elif k == 1:
            out += tag + b"Z" + bytes(rng.choice(np.frombuffer(b"NM:i", dtype=np.uint8), int(rng.integers(0, 12)))) + b"\0"
        elif k == 2:
            out += tag + b"f" + struct.pack("<f", 1.5)
        elif k == 3:
            out += tag + b"d" + struct.pack("<d", 2.5)
        elif k == 4:
            cnt = int(rng.integers(0, 6))
            sub = [(b"c", 1), (b"C", 1), (b"s", 2), (b"S", 2), (b"i", 4), (b"I", 4), (b"f", 4)][int(rng.integers(0, 7))]
            out += tag + b"B" + sub[0] + struct.pack("<I", cnt) + bytes(rng.integers(0, 256, cnt * sub[1], dtype=np.uint8))
        elif k == 5:
            out += tag + b"H" + b"4E4D" + b"\0"
        elif k == 6:
            out += tag + b"S" + struct.pack("<H", 513)
        else:
            out += tag + b"i" + struct.pack("<i", -7)
    k = rng.random()
    v = int(rng.choice([0, 0, 0, 1, 2, 3, 4, 9]))
    if k < 0.15:
        pass                                           # no NM
    elif k < 0.35:
        out += b"NMC" + bytes([v])
    elif k < 0.45:
        out += b"NMc" + struct.pack("<b", v if rng.random() < 0.8 else -1)
    elif k < 0.55:
        out += b"NMs" + struct.pack("<h", v)
    elif k < 0.65:
        out += b"NMS" + struct.pack("<H", v)
    elif k < 0.8:
        out += b"NMi" + struct.pack("<i", v)
    elif k < 0.9:
        out += b"NMI" + struct.pack("<I", v)
    else:
        out += b"NMZ" + b"%d\0" % v                    # bam_aux2i of a string tag is 0
    if rng.random() < 0.7:
        out += b"RGZgrp%d\0" % int(rng.integers(0, 3))
    return out


def _cigar(rng, rl):
    k = rng.random()
    if rl < 12 or k < 0.4:
        return [(0, rl)] if rl else []
    if k < 0.45:
        return []                                      # mapped without a CIGAR
    if k < 0.55:
        return [(7, rl)] if rng.random() < 0.5 else [(7, rl - 5), (8, 1), (7, 4)]       # = and X
    a = int(rng.integers(3, rl - 6))
    if k < 0.65:
        return [(4, a), (0, rl - a)]
    if k < 0.72:
        return [(5, 7), (0, rl - a), (4, a)]
    if k < 0.8:
        return [(0, a), (2, int(rng.integers(1, 60))), (0, rl - a)]
    if k < 0.86:
        return [(0, a), (1, 3), (0, rl - a - 3)]
    if k < 0.92:
        return [(0, a), (3, int(rng.integers(50, 3000))), (0, rl - a)]
    if k < 0.96:
        return [(0, a), (6, 2), (0, rl - a)]
    return [(0, a), (2, 1)]                            # two operations with a deletion: not an indel for HasIndel


def _sequence(rng, rl):
    s = bytearray(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), rl).tobytes())
    k = rng.random()
    if rl and k < 0.25:
        a, b = int(rng.integers(0, min(rl, 9))), int(rng.integers(0, min(rl, 9)))
        s[:a] = b"N" * a
        s[rl - b:] = b"N" * b
    elif rl and k < 0.4:
        for p in rng.integers(0, rl, int(rng.integers(1, max(2, rl // 6)))):
            s[p] = ord("N")
    elif rl and k < 0.5:
        a, b = int(rng.integers(0, min(rl, 4))), int(rng.integers(0, min(rl, 4)))
        s[:a] = rng.choice(np.frombuffer(b"=MRSVWYHKDB", dtype=np.uint8), a).tobytes()
        s[rl - b:] = rng.choice(np.frombuffer(b"=MRSVWYHKDB", dtype=np.uint8), b).tobytes()
    elif k < 0.53:
        s[:] = b"N" * rl
    return bytes(s)


def fuzz_stream(seed: int, n_names: int = 1500, refs=REFS) -> bytes:
    rng = np.random.default_rng(seed)
    recs = []
    stem = b"instrument:run:flowcell:" + b"x" * int(rng.integers(0, 200))
    for i in range(n_names):
        name = (stem if rng.random() < 0.5 else b"r") + b"%d" % i
        name = name[-254:]
        tid = int(rng.integers(0, len(refs)))
        base = int(rng.integers(0, refs[tid][1] - 5000))
        for occ in range(int(rng.choice([1, 2, 2, 2, 2, 2, 3, 4, 5]))):
            rl = int(rng.choice([0, 15, 21, 22, 23, 36, 76, 101, 151, 250]))
            pos = base + int(rng.integers(0, 1200))
            flag = 1 | int(rng.choice([0x40, 0x80, 0xC0, 0])) | (16 if rng.random() < 0.5 else 0) | (32 if rng.random() < 0.5 else 0)
            flag |= int(rng.choice([0, 0, 0, 0, 0, 0x100, 0x200, 0x400, 0x800, 2, 8]))
            cig = _cigar(rng, rl)
            if rng.random() < 0.12:
                flag |= 4
                cig = [] if rng.random() < 0.7 else cig
            mapq = int(rng.choice([0, 1, 10, 29, 30, 31, 60, 255]))
            nm = name + b"\0"
            seq = _sequence(rng, rl)
            body = struct.pack("<iiBBHHHiiii", tid, pos, len(nm), mapq, 4680, len(cig), flag, rl, tid, base, 0) + nm + \
                b"".join(struct.pack("<I", (ln << 4) | op) for op, ln in cig) + pack_seq(seq) + bytes(rng.integers(0, 42, rl, dtype=np.uint8)) + _aux(rng)
            recs.append((tid, pos, len(recs), struct.pack("<I", len(body)) + body))
    recs.sort(key=lambda r: (r[0], r[1], r[2]))
    return bam.header(list(refs)) + b"".join(r[3] for r in recs)


QUERIES = ((0, 0, 200_000, 800, 0), (1, 0, 150_000, 800, 30), (0, 50_000, 120_000, 300, 10), (1, 149_000, 150_000, 800, 0))
