"""Minimal BAM writer / serial reader used by the tests and the benchmark of the device record index (SAM specification
section 4.2): enough of the format to produce valid alignment records with names, CIGARs, sequences, qualities and tags."""
from __future__ import annotations

import struct

import numpy as np

REC_DT = np.dtype([("off", "<u8"), ("tid", "<i4"), ("pos", "<i4"), ("end", "<i4"), ("flag", "<u2"), ("mapq", "u1"), ("pad", "u1")])
# pg_sr_query / pg_sr_cand / pg_sr_ref (include/pindel_pg.h)
SR_QUERY_DT = np.dtype([("tid", "<i4"), ("beg", "<i4"), ("end", "<i4"), ("insert_size", "<i4"), ("min_anchor_quality", "<u4"), ("chr_len", "<u4"),
                        ("nm_limit", "<i4"), ("pad", "<u4"), ("max_mismatch_rate", "<f8")])
SR_CAND_DT = np.dtype([("anchor", "<u4"), ("read", "<u4"), ("matched_rel_pos", "<u4"), ("read_length", "<u4"), ("matched_d", "u1"), ("ms", "u1"), ("pad", "<u2")])
SR_REF_DT = np.dtype([("pos", "<u4"), ("read_length", "<i2"), ("mq", "u1"), ("pad", "u1")])


def header(refs, text: bytes = b"@HD\tVN:1.6\tSO:coordinate\n") -> bytes:
    out = [b"BAM\x01", struct.pack("<i", len(text)), text, struct.pack("<i", len(refs))]
    for name, length in refs:
        nm = name.encode() + b"\0"
        out += [struct.pack("<i", len(nm)), nm, struct.pack("<i", length)]
    return b"".join(out)


def record(tid, pos, name: bytes, flag, mapq, cigar, seq_len, mtid=-1, mpos=-1, tlen=0, aux: bytes = b"", rng=None) -> bytes:
    """cigar: list of (op, len) with op in 0..8 (MIDNSHP=X)."""
    nm = name + b"\0"
    cig = b"".join(struct.pack("<I", (ln << 4) | op) for op, ln in cigar)
    rng = rng or np.random.default_rng(0)
    seq = rng.integers(0, 256, (seq_len + 1) // 2, dtype=np.uint8).tobytes()
    qual = rng.integers(0, 42, seq_len, dtype=np.uint8).tobytes()
    body = struct.pack("<iiBBHHHiiii", tid, pos, len(nm), mapq, 4680, len(cigar), flag, seq_len, mtid, mpos, tlen) + nm + cig + seq + qual + aux
    return struct.pack("<I", len(body)) + body


def first_record_offset(stream: bytes):
    """-> (offset of the first alignment record, n_ref)"""
    p = 8 + struct.unpack_from("<i", stream, 4)[0]
    n_ref = struct.unpack_from("<i", stream, p)[0]
    p += 4
    for _ in range(n_ref):
        p += 4 + struct.unpack_from("<i", stream, p)[0] + 4
    return p, n_ref


def serial_index(stream: bytes) -> np.ndarray:
    """The record table by the plain serial walk (what oracle/shim/minihts.cpp and htslib do)."""
    p, _ = first_record_offset(stream)
    rows = []
    n = len(stream)
    while p + 4 <= n:
        bs = struct.unpack_from("<I", stream, p)[0]
        tid, pos, l_name, mapq, _bin, n_cig, flag = struct.unpack_from("<iiBBHHH", stream, p + 4)
        span = 0
        if not flag & 4:
            for k in range(n_cig):
                c = struct.unpack_from("<I", stream, p + 36 + l_name + 4 * k)[0]
                if c & 15 in (0, 2, 3, 7, 8):
                    span += c >> 4
        rows.append((p, tid, pos, pos + (span if span > 0 else 1), flag, mapq, 0))
        p += 4 + bs
    assert p == n, "truncated BAM record"
    return np.array(rows, dtype=REC_DT) if rows else np.zeros(0, dtype=REC_DT)


def synthetic_stream(n_records: int, seed: int, refs=(("chr1", 50_000_000), ("chr2", 30_000_000)), adversarial: bool = False,
                     discordant: float = 0.0, switch_ref: bool = True, unplaced: float = 0.02) -> bytes:
    """A coordinate-sorted BAM stream (inflated).  adversarial: sequence / quality / tag bytes that contain whole fake
    records (valid-looking headers with names), so that a segment boundary falling into them makes the guess lock on."""
    rng = np.random.default_rng(seed)
    out = [header(list(refs))]
    pos = 1000
    tid = 0
    fake = record(0, 12345, b"fake_read_inside_aux", 99, 60, [(0, 100)], 100, 0, 12600, 350, rng=rng)
    for i in range(n_records):
        pos += int(rng.integers(0, 60))
        if switch_ref and i == n_records // 2:
            tid, pos = 1, 500
        rl = int(rng.choice([76, 100, 150, 151]))
        kind = rng.random()
        if kind < 0.7:
            cig = [(0, rl)]
        elif kind < 0.8:
            a = int(rng.integers(10, rl - 10)); cig = [(4, a), (0, rl - a)]
        elif kind < 0.9:
            a = int(rng.integers(10, rl - 20)); cig = [(0, a), (2, int(rng.integers(1, 50))), (0, rl - a)]
        elif kind < 0.95:
            a = int(rng.integers(10, rl - 20)); cig = [(0, a), (1, 5), (0, rl - a - 5)]
        else:
            cig = []
        flag = 4 | 69 if not cig else int(rng.choice([99, 147, 83, 163, 65, 129]))
        aux = b"NMC" + bytes([int(rng.integers(0, 5))]) + b"RGZgrp%d\0" % int(rng.integers(0, 3))
        if adversarial and rng.random() < 0.3:
            aux += b"XZB" + b"C" + struct.pack("<I", 3 * len(fake)) + fake * 3       # B-array tag carrying three fake records
        mtid, mpos, tlen = tid, pos + int(rng.integers(100, 600)), int(rng.integers(-600, 600))
        if rng.random() < discordant:          # pairs the discovery keeps: far mate, mate on the other reference, mate before the read
            k = rng.random()
            if k < 0.4:
                tlen = int(rng.integers(2000, 900_000)) * (1 if rng.random() < 0.5 else -1)
                mpos = max(0, pos + tlen)
            elif k < 0.7:
                mtid = (1 - tid) if len(refs) > 1 else tid
                mpos = int(rng.integers(0, 1_000_000))
            else:
                flag = int(rng.choice([65, 129, 113, 177]))         # both mates on one strand
                mpos = max(0, pos - int(rng.integers(0, 3000)))
            if rng.random() < 0.1:
                flag |= 8                                           # mate unmapped: not a pair for the discovery
        out.append(record(tid if cig else -1 if rng.random() < unplaced else tid, pos, b"read_%d_%d" % (seed, i), flag, int(rng.integers(0, 61)), cig, rl,
                          mtid, mpos, tlen, aux, rng))
    return b"".join(out)
