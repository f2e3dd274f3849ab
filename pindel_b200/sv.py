"""Host side of the SV-classification stage (include/pindel_pg.h, pg_classify): input marshalling and the replay of the
records the device returns -- what a Pindel host does with them per driver: write the SPLIT_READ fields
(src/pindel.h:349-360), Vars[box].push_back(ReadIndex), FutureReads_SR.push_back(read) (src/pindel.cpp:587-592).

Reference drivers: src/search_variant.cpp:49-266, src/search_deletions_nt.cpp:26, src/search_tandem_duplications.cpp:28,
src/search_tandem_duplications_nt.cpp:27, src/search_inversions.cpp:29, src/search_inversions_nt.cpp:27.
"""
from __future__ import annotations

import numpy as np

from . import runs as R

DRIVERS = ("D", "DI", "TD", "TD_NT", "INV", "INV_NT", "SI")
ALL_DRIVERS = 0x7F
NONE, BOX, FUTURE, BREAK = 0, 1, 2, 3
F_CORE, F_NTSTR, F_NTSIZE, F_KILLED = 1, 2, 4, 8
NT_EMPTY, NT_SEQ, NT_RC = 0, 1, 2
R_USED, R_OTHERCHR = 1, 2

RECORD_DT = np.dtype([("read", "<u4"), ("driver", "u1"), ("outcome", "u1"), ("fields", "u1"), ("seq", "u1"),
                      ("left", "<i4"), ("right", "<i4"), ("bp_left", "<u4"), ("bp_right", "<u4"), ("indel_size", "<u4"),
                      ("bp", "<i2"), ("nt_size", "<u2"), ("nt_start", "<u2"), ("nt_len", "<u2"), ("nt_rot", "<u2"),
                      ("nt_src", "u1"), ("pad", "u1"), ("box", "<u4")])
assert RECORD_DT.itemsize == 44


def window(chr_comp_size: int, window_end: int, chr_id: int = 0, driver_mask: int = ALL_DRIVERS, region=None,
           min_inversion_size: int = 30, min_matched_bases: int = 30, box_size=None, num_boxes=None) -> dict:
    """Window parameters as main() derives them: BoxSize = compSize / 30000 (at least 1), NumBoxes = compSize * 2 /
    BoxSize + 1 (src/pindel.cpp:1806-1810); region = (start, end) or None for "not defined"."""
    bs = box_size if box_size is not None else (chr_comp_size // 30000 or 1)
    nb = num_boxes if num_boxes is not None else chr_comp_size * 2 // bs + 1
    return dict(driver_mask=driver_mask, chr=chr_id, window_end=window_end, num_boxes=nb, box_size=bs, region=region,
                min_inversion_size=min_inversion_size, min_matched_bases=min_matched_bases)


def compress_ups(rows: np.ndarray) -> np.ndarray:
    """[k, >=5] UniquePoint rows (len, absLoc, dir, strand, mm[, chr]) -> runs (RUN_DT); inverse of runs.expand_runs."""
    k = rows.shape[0]
    if k == 0:
        return np.zeros(0, dtype=R.RUN_DT)
    ln, loc, d, s, mm = (rows[:, i].astype(np.int64) for i in range(5))
    ch = rows[:, 5].astype(np.int64) if rows.shape[1] > 5 else np.zeros(k, np.int64)
    loc &= 0xFFFFFFFF
    step = np.where(d[:-1] == 0, 1, -1)
    cont = ((ln[1:] == ln[:-1] + 1) & (((loc[:-1] + step) & 0xFFFFFFFF) == loc[1:]) & (d[1:] == d[:-1]) & (s[1:] == s[:-1])
            & (mm[1:] == mm[:-1]) & (ch[1:] == ch[:-1]))
    starts = np.concatenate(([0], np.nonzero(~cont)[0] + 1))
    ends = np.concatenate((starts[1:], [k]))
    out = np.zeros(starts.size, dtype=R.RUN_DT)
    out["abs_loc0"] = loc[starts]
    out["len0"] = ln[starts]
    out["count"] = ends - starts
    out["mm"] = mm[starts]
    out["dirstrand"] = d[starts] | (s[starts] << 1)
    out["chr"] = ch[starts]
    return out


def batch_from_lists(seqs, matched_d, insert_size, close_lists, far_lists, rflags=None) -> dict:
    """seqs: final UnmatchedSeq per read (bytes); close_lists / far_lists: per-read [k, >=5] rows (UP_Close after
    CleanUniquePoints, UP_Far).  -> the arrays of pg_sv_reads."""
    n = len(seqs)
    runs, run_off, n_close = [], [0], []
    for c, f in zip(close_lists, far_lists):
        rc, rf = compress_ups(np.asarray(c)), compress_ups(np.asarray(f))
        runs += [rc, rf]
        n_close.append(rc.size)
        run_off.append(run_off[-1] + rc.size + rf.size)
    off = np.zeros(n + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(s) for s in seqs])
    return dict(seq_blob=np.frombuffer(b"".join(seqs), dtype=np.uint8), seq_off=off,
                matched_d=np.ascontiguousarray(matched_d, dtype=np.uint8), insert_size=np.ascontiguousarray(insert_size, dtype=np.int16),
                rflags=np.zeros(n, np.uint8) if rflags is None else np.ascontiguousarray(rflags, dtype=np.uint8),
                run_off=np.array(run_off, dtype=np.uint32), n_close_runs=np.array(n_close, dtype=np.uint16),
                runs=np.concatenate(runs) if runs else np.zeros(0, dtype=R.RUN_DT))


def nt_string(seq: bytes, rec) -> bytes:
    """NT_str of a record: rotl(substr of the read or of its reverse complement)."""
    if rec["nt_src"] == NT_EMPTY or rec["nt_len"] == 0:
        return b""
    src = seq if rec["nt_src"] == NT_SEQ else R.reverse_complement(seq)
    s = src[int(rec["nt_start"]):int(rec["nt_start"]) + int(rec["nt_len"])]
    k = int(rec["nt_rot"])
    return s[k:] + s[:k]


class ReadState:
    """The SPLIT_READ fields the drivers write."""
    __slots__ = ("Left", "Right", "BP", "BPLeft", "BPRight", "IndelSize", "NT_size", "NT_str", "Used")

    def __init__(self, used=False):
        self.Left = self.Right = self.BP = self.BPLeft = self.BPRight = self.IndelSize = self.NT_size = 0
        self.NT_str = b""
        self.Used = bool(used)

    def copy(self):
        c = ReadState()
        for k in self.__slots__:
            setattr(c, k, getattr(self, k))
        return c


def replay(records: np.ndarray, seqs, rflags, n_boxes_hint=None):
    """Apply records (sorted by driver, read, seq -- as pg_classify returns them) the way the reference's drivers act.
    -> (states [n] ReadState after all drivers, per_driver {d: dict(vars=[(box, read)...] box-major, future=[(read, ReadState)...],
    state=[n] snapshot after driver d)})."""
    n = len(seqs)
    st = [ReadState(bool(rflags[i] & R_USED)) for i in range(n)]
    per = {}
    drivers = sorted(set(int(d) for d in records["driver"])) if records.size else []
    for d in drivers:
        recs = records[records["driver"] == d]
        vars_, future = [], []
        for rec in recs:
            i = int(rec["read"])
            s = st[i]
            f = int(rec["fields"])
            if f & F_CORE:
                s.Left, s.Right, s.BP = int(rec["left"]), int(rec["right"]), int(rec["bp"])
                s.BPLeft, s.BPRight, s.IndelSize = int(rec["bp_left"]), int(rec["bp_right"]), int(rec["indel_size"])
            if f & F_NTSTR:
                s.NT_str = nt_string(seqs[i], rec)
            if f & F_NTSIZE:
                s.NT_size = int(rec["nt_size"])
            if f & F_KILLED:
                s.Used = True
            o = int(rec["outcome"])
            if o == FUTURE:
                future.append((i, s.copy()))
                s.Used = True
            elif o == BOX:
                vars_.append((int(rec["box"]), i))
                s.Used = True
            elif o == BREAK:
                s.Used = True
        vars_.sort(key=lambda t: t[0])  # stable: push order inside a box is read order
        per[d] = dict(vars=vars_, future=future, state=[s.copy() for s in st])
    return st, per
