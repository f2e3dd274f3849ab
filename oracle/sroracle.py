"""oracle/sroracle.py -- TEST INFRASTRUCTURE ONLY: plain-Python restatement of what ReadInBamReads_SR (src/reader.cpp:483-560)
hands to the ReadBuffer for one BAM file and window -- the groundwork (oracle first) for the rest of SURVEY.md section 8(f) row
N3, the reader's split-read filter:

    fetch_func_SR            src/reader.cpp:1099-1151   mate pairing on the read name, as the records stream by
    parse_flags_and_tags     src/reader.cpp:1258-1316   mapped, NM
    isGoodAnchor / isRefRead / isWeirdRead               src/reader.cpp:561-690
    build_record_SR          src/reader.cpp:799-900     sequence decode, N trimming, reverse complement, anchor position
    build_record_RefRead     src/reader.cpp:903-923

Pinned against the reference's own function by tests/test_sr_oracle_pinning.py (committed dump tests/golden/sr_*.npz made
through oracle/ref_hook.cpp pgref_sr_candidates, and live when oracle/_ref is built).  No device code consumes this yet.
"""
from __future__ import annotations

import struct

import numpy as np

from pindel_b200 import runs as R

SEQ16 = b"=ACMGRSVTWYHKDBN"
SPACER = 100000


def fnv1a(b: bytes) -> int:
    h = 2166136261
    for c in b:
        h = ((h ^ c) * 16777619) & 0xFFFFFFFF
    return h


class Rec:
    """The fields of a bam1_t the reader looks at."""
    __slots__ = ("tid", "pos", "qual", "flag", "l_qseq", "mtid", "mpos", "isize", "qname", "cigar", "seq", "aux")

    def __init__(self, stream: bytes, off: int):
        bs = struct.unpack_from("<I", stream, off)[0]
        self.tid, self.pos, l_name, self.qual, _bin, n_cig, self.flag, self.l_qseq, self.mtid, self.mpos, self.isize = \
            struct.unpack_from("<iiBBHHHiiii", stream, off + 4)
        p = off + 36
        self.qname = stream[p:p + l_name].split(b"\0")[0]
        p += l_name
        self.cigar = [struct.unpack_from("<I", stream, p + 4 * k)[0] for k in range(n_cig)]
        p += 4 * n_cig
        self.seq = stream[p:p + (self.l_qseq + 1) // 2]
        p += (self.l_qseq + 1) // 2 + self.l_qseq
        self.aux = stream[p:off + 4 + bs]


_AUX_SIZE = {ord("A"): 1, ord("c"): 1, ord("C"): 1, ord("s"): 2, ord("S"): 2, ord("i"): 4, ord("I"): 4, ord("f"): 4, ord("d"): 8}


def aux_get(aux: bytes, tag: bytes):
    """-> (type, value bytes) of the first occurrence of `tag`, or None (bam_aux_get)"""
    p = 0
    while p + 3 <= len(aux):
        t, ty = aux[p:p + 2], aux[p + 2]
        p += 3
        if ty in _AUX_SIZE:
            n = _AUX_SIZE[ty]
        elif ty in (ord("Z"), ord("H")):
            n = aux.index(b"\0", p) - p + 1
        elif ty == ord("B"):
            sub, cnt = aux[p], struct.unpack_from("<I", aux, p + 1)[0]
            n = 5 + _AUX_SIZE[sub] * cnt
        else:
            return None
        if t == tag:
            return ty, aux[p:p + n]
        p += n
    return None


def aux2i(v):
    ty, b = v
    fmt = {ord("c"): "<b", ord("C"): "<B", ord("s"): "<h", ord("S"): "<H", ord("i"): "<i", ord("I"): "<I"}.get(ty)
    return struct.unpack(fmt, b)[0] if fmt else 0


def cigar2mismatch(r: Rec) -> int:      # src/reader.cpp:1333-1345: every base of a non-M operation
    return sum(c >> 4 for c in r.cigar if c & 15 != 0)


def cigar2len(r: Rec) -> int:           # src/reader.cpp:1318-1331: M, I, S add, D subtracts
    n = 0
    for c in r.cigar:
        op = c & 15
        if op in (0, 1, 4):
            n += c >> 4
        if op == 2:
            n -= c >> 4
    return n


def flags_of(r: Rec):
    nm = aux_get(r.aux, b"NM")
    return {"mapped": not r.flag & 4, "edits": aux2i(nm) if nm else 0}


def is_good_anchor(fl, r: Rec, min_q: int) -> bool:
    if not fl["mapped"] or r.qual < min_q:
        return False
    if min_q == 0:
        return True
    return not r.flag & (0x100 | 0x200 | 0x400)


def is_weird_read(fl, r: Rec) -> bool:
    if not fl["mapped"]:
        return True
    if any(c & 15 in (1, 2, 3, 4, 5, 6) for c in r.cigar):
        return True
    nm = aux_get(r.aux, b"NM")
    if nm and aux2i(nm):
        return True
    return fl["edits"] + cigar2mismatch(r) > 0


def is_ref_read(fl, r: Rec, max_mismatch_rate: float, nm_limit: int) -> bool:
    if r.flag & (0x100 | 0x200 | 0x400):
        return False
    nm = aux_get(r.aux, b"NM")
    if nm:
        v = aux2i(nm)
        if v > nm_limit or v > int(r.l_qseq * max_mismatch_rate) + 1:
            return False
    if len(r.cigar) > 2 and any(c & 15 in (1, 2) for c in r.cigar):      # HasIndel
        return False
    return fl["mapped"] and fl["edits"] <= 2 and cigar2mismatch(r) <= 2


def build_record_sr(mapped: Rec, unmapped: Rec, insert_size: int, chr_len: int, min_q: int):
    """-> (Name, UnmatchedSeq, MatchedD, MatchedRelPos, MS) or None"""
    if mapped.qual < min_q:
        return None
    name = b"@" + unmapped.qname + (b"/1" if unmapped.flag & 0x40 else b"/2" if unmapped.flag & 0x80 else b"")
    seq = bytes(SEQ16[(unmapped.seq[i >> 1] >> (0 if i & 1 else 4)) & 15] for i in range(unmapped.l_qseq))
    seq = seq.lstrip(b"N")
    length = len(seq)
    seq = seq.rstrip(b"N")
    length = len(seq)
    if seq.count(b"N") > int(length * 0.10) or length < 22:
        return None
    s = R.set_unmatched_seq(R.reverse_complement(seq) if unmapped.flag & 16 else seq)
    pos = mapped.pos
    if mapped.flag & 16:
        d = ord("-")
        pos += cigar2len(mapped)
    else:
        d = ord("+")
    pos &= 0xFFFFFFFF
    if insert_size <= length:
        raise ValueError("insert size not larger than the read length (the reference exits here)")
    if pos > chr_len:
        pos = chr_len
    return name, s, d, pos, mapped.qual


def candidates(stream: bytes, table: np.ndarray, tid: int, beg: int, end: int, insert_size: int, chr_len: int, min_anchor_quality: int = 0,
               max_mismatch_rate: float = 0.02, nm_limit: int = 2):
    """table: pindel_b200.bam.serial_index(stream).  -> (candidates [k, 7] like oracle.refhook.sr_candidates, ref reads [m, 3])"""
    waiting = {}
    cand, refs = [], []
    beg = max(beg, 0)

    def sr(m, u):
        r = build_record_sr(m, u, insert_size, chr_len, min_anchor_quality)
        if r is not None:
            name, s, d, pos, ms = r
            cand.append((fnv1a(name), d, pos, ms, insert_size, len(s), fnv1a(s)))

    def ref(m, rr):
        if rr.qual >= min_anchor_quality:
            refs.append((rr.pos & 0xFFFFFFFF, rr.qual, ((rr.l_qseq + 0x8000) & 0xFFFF) - 0x8000))

    for row in table:
        rt, pos = int(row["tid"]), int(row["pos"])
        if rt != tid:
            if rt > tid or rt < 0:
                break
            continue
        if pos >= end:
            break
        if not int(row["end"]) > beg:
            continue
        b1 = Rec(stream, int(row["off"]))
        f1 = flags_of(b1)
        b2 = waiting.pop(b1.qname, None)
        if b2 is None:
            waiting[b1.qname] = b1
            if is_weird_read(f1, b1):
                sr(b1, b1)
            continue
        f2 = flags_of(b2)
        if is_weird_read(f2, b2):
            sr(b2, b2)
        if is_good_anchor(f1, b1, min_anchor_quality):
            if is_weird_read(f2, b2):
                sr(b1, b2)
            if is_ref_read(f2, b2, max_mismatch_rate, nm_limit):
                ref(b1, b2)
        if is_good_anchor(f2, b2, min_anchor_quality):
            if is_weird_read(f1, b1):
                sr(b2, b1)
            if is_ref_read(f1, b1, max_mismatch_rate, nm_limit):
                ref(b2, b1)
    return np.array(cand, dtype=np.int64).reshape(-1, 7), np.array(refs, dtype=np.int64).reshape(-1, 3)
