"""oracle/rpdoracle.py -- TEST INFRASTRUCTURE ONLY: plain-Python restatement of the discordant-pair discovery for one BAM
file and window -- the region iterator as the reference sees it through oracle/shim/minihts.cpp (sam_itr_next) and
build_record_RP_Discovery (src/reader.cpp:982-1097).  Pinned against the reference's own ReadInBamReads_RP_Discovery by
tests/test_rp_discovery.py (committed dumps tests/golden/rpd_*.npz, and live when oracle/_ref is built)."""
from __future__ import annotations

import struct

import numpy as np


def fnv1a(b: bytes) -> int:
    h = 2166136261
    for c in b:
        h = ((h ^ c) * 16777619) & 0xFFFFFFFF
    return h


def discovery(stream: bytes, table: np.ndarray, tid: int, beg: int, end: int, insert_size: int, min_anchor_quality: int = 0):
    """table: the serial record table (pindel_b200.bam.serial_index).  -> (rows [k, 9] like oracle.refhook.rp_discovery:
    list, PosA, PosB, DA, DB, ReadLength, InsertSize, MQA, FNV-1a of the read name; src record indices; records delivered)."""
    rows, src, scanned = [], [], 0
    beg = max(beg, 0)
    for i, r in enumerate(table):
        rt, pos = int(r["tid"]), int(r["pos"])
        if rt != tid:
            if rt > tid or rt < 0:
                break                                        # sam_itr_next: past the chromosome
            continue
        if pos >= end:
            break
        if not int(r["end"]) > beg:
            continue
        scanned += 1
        off = int(r["off"]) + 4
        _tid, _pos, l_name, qual, _bin, _ncig, flag, l_seq, mtid, mpos, isize = struct.unpack_from("<iiBBHHHiiii", stream, off)
        if not flag & 1 or qual < min_anchor_quality or flag & 4 or flag & 8:
            continue
        rev, mrev = bool(flag & 16), bool(flag & 32)
        if not (rt != mtid or abs(isize) > 3 * insert_size + 1000 or rev == mrev):
            continue
        name = stream[off + 32:off + 32 + l_name - 1]
        da, db = (45 if rev else 43), (45 if mrev else 43)
        pa, pb = pos & 0xFFFFFFFF, mpos & 0xFFFFFFFF
        rl = ((l_seq + 0x8000) & 0xFFFF) - 0x8000          # short ReadLength
        if rt == mtid:
            if not pa < pb:
                pa, pb, da, db = pb, pa, db, da
            rows.append((0, pa, pb, da, db, rl, insert_size, qual, fnv1a(name)))
        else:
            rows.append((1, pa, pb, da, db, rl, insert_size, qual, fnv1a(name)))
        src.append(i)
    out = np.array(rows, dtype=np.int64).reshape(-1, 9)
    # the reference keeps two vectors: list 0 rows first, then list 1, each in BAM order
    order = np.argsort(out[:, 0], kind="stable") if out.size else np.zeros(0, dtype=np.int64)
    return out[order], np.array(src, dtype=np.int64)[order], scanned
