"""Host-side view of the search results: run-length coded UniquePoints.

The kernels emit, per read, runs of consecutive UniquePoints (src/pindel.h:137-158) that share the
seed and the mismatch count; this module expands them into the reference's flat
(LengthStr, AbsLoc, Direction, Strand, Mismatches, chromosome) records and mirrors
CleanUniquePoints (src/pindel.cpp:2904-2941), which the reference applies to UP_Close right after
GetCloseEnd (src/reader.cpp:272, :1541).
"""
from __future__ import annotations

import numpy as np

RUN_DT = np.dtype([("abs_loc0", "<u4"), ("len0", "<u2"), ("count", "<u2"), ("mm", "u1"), ("dirstrand", "u1"),
                   ("chr", "<i2")])
OUT_DT = np.dtype([("run_off", "<u4"), ("n_close_runs", "<u2"), ("n_far_runs", "<u2"), ("len_after", "<u2"),
                   ("flags", "<u2")])
REGION_DT = np.dtype([("chr", "<i4"), ("start", "<u4"), ("end", "<u4")])
CLOSEINFO_DT = np.dtype([("last_abs_loc", "<u4"), ("max_len", "<u2"), ("pad", "<u2")])

F_FLIPPED, F_TOOLONG, F_OVERFLOW = 1, 2, 4
assert RUN_DT.itemsize == 12 and OUT_DT.itemsize == 12 and REGION_DT.itemsize == 12 and CLOSEINFO_DT.itemsize == 8


def expand_runs(runs: np.ndarray) -> np.ndarray:
    """runs (RUN_DT) -> int64 [k, 6] rows (len, absLoc, dir, strand, mm, chr) in emission order."""
    if runs.size == 0:
        return np.zeros((0, 6), dtype=np.int64)
    cnt = runs["count"].astype(np.int64)
    tot = int(cnt.sum())
    rid = np.repeat(np.arange(runs.size), cnt)
    first = np.cumsum(cnt) - cnt
    j = np.arange(tot) - first[rid]
    r = runs[rid]
    d = (r["dirstrand"] & 1).astype(np.int64)
    out = np.empty((tot, 6), dtype=np.int64)
    out[:, 0] = r["len0"].astype(np.int64) + j
    out[:, 1] = (r["abs_loc0"].astype(np.int64) + np.where(d == 0, j, -j)) & 0xFFFFFFFF
    out[:, 2] = d
    out[:, 3] = (r["dirstrand"] >> 1) & 1
    out[:, 4] = r["mm"]
    out[:, 5] = r["chr"]
    return out


def clean_unique_points(ups: np.ndarray) -> np.ndarray:
    """CleanUniquePoints (src/pindel.cpp:2904-2941) on [k, 6] rows."""
    if ups.shape[0] == 0:
        return ups
    last = ups[-1]
    if last[2] == 0:
        term = (ups[:, 1] - ups[:, 0]) & 0xFFFFFFFF
        t0 = (last[1] - last[0]) & 0xFFFFFFFF
    else:
        term = (ups[:, 1] + ups[:, 0]) & 0xFFFFFFFF
        t0 = (last[1] + last[0]) & 0xFFFFFFFF
    keep = (ups[:, 5] == last[5]) & (ups[:, 2] == last[2]) & (ups[:, 3] == last[3]) & (term == t0)
    return ups[keep]


_RC = bytes.maketrans(b"ACGTN", b"TGCAN")
_KEEP = set(b"ACGTN")


def set_unmatched_seq(seq: bytes) -> bytes:
    """SPLIT_READ::setUnmatchedSeq (src/pindel.cpp:142-169): trailing non-alphanumerics are stripped."""
    n = len(seq)
    while n > 0 and not chr(seq[n - 1]).isalnum():
        n -= 1
    return seq[:n]


def reverse_complement(seq: bytes) -> bytes:
    """ReverseComplement (src/pindel.cpp:2037-2048, Convert2RC4N): anything but A/C/G/T/N becomes NUL."""
    return bytes((c if c in _KEEP else 0) for c in seq[::-1]).translate(_RC)


def final_sequence(seq: bytes, nflips: int) -> bytes:
    """The read's UnmatchedSeq after GetCloseEnd: the host replays the flips the device reports (PG_READ_NFLIPS)."""
    s = set_unmatched_seq(seq)
    for _ in range(nflips):
        s = set_unmatched_seq(reverse_complement(s))
    return s


def per_read(out: np.ndarray, runs: np.ndarray, seq_blob=None, seq_off=None):
    """-> list of dicts {flip, L, close_raw, close, far, flags, nflips} (same keys as oracle.refhook.parse_stream).
    With the input sequences given, `flip` means what it means in the oracle streams: the final UnmatchedSeq differs
    from the input; without them it is the parity of the flips."""
    res = []
    for i, o in enumerate(out):
        off, nc, nf = int(o["run_off"]), int(o["n_close_runs"]), int(o["n_far_runs"])
        cr = expand_runs(runs[off:off + nc])
        fa = expand_runs(runs[off + nc:off + nc + nf])
        flags = int(o["flags"])
        nflips = (flags >> 8) & 3
        flip = flags & F_FLIPPED
        if seq_blob is not None:
            orig = bytes(seq_blob[int(seq_off[i]):int(seq_off[i + 1])])
            fin = final_sequence(orig, nflips)
            assert len(fin) == int(o["len_after"]) or (flags & F_TOOLONG), (i, len(fin), int(o["len_after"]))
            flip = int(fin != orig)
        res.append(dict(flip=flip, L=int(o["len_after"]), close_raw=cr, close=clean_unique_points(cr), far=fa,
                        flags=flags, nflips=nflips))
    return res
