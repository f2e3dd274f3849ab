"""ctypes wrapper of oracle/boxsort_oracle.c (TEST INFRASTRUCTURE ONLY) and of the reference's own bubblesortReads /
markDuplicates through oracle/ref_hook.cpp, plus the synthetic box generator the tests use."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

KEY_DT = np.dtype([("bp_left", "<u4"), ("bp_right", "<u4"), ("indel_size", "<u4"), ("left", "<i4"), ("right", "<i4"), ("name_id", "<u4"),
                   ("nt_size", "<u2"), ("bp", "<i2"), ("unique", "u1"), ("matched_d", "u1"), ("read_length", "<u2")])
assert KEY_DT.itemsize == 32


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libboxsort_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/libboxsort_oracle.so missing: run `make -C oracle port`")
        _lib = C.CDLL(path)
    return _lib


def sort_boxes(reads: np.ndarray, box_off: np.ndarray, idx: np.ndarray):
    """reads: KEY_DT per read; idx: read indices grouped by box.  -> (idx reordered per box, UniqueRead per read)"""
    reads = np.ascontiguousarray(reads, dtype=KEY_DT)
    idx = np.ascontiguousarray(idx, dtype=np.uint32).copy()
    unique = reads["unique"].copy()
    for b in range(len(box_off) - 1):
        a, e = int(box_off[b]), int(box_off[b + 1])
        lib().bso_box(reads.ctypes.data_as(C.c_void_p), unique.ctypes.data_as(C.c_void_p), C.c_void_p(idx.ctypes.data + 4 * a), C.c_uint32(e - a))
    return idx, unique


def sort_boxes_kind(kind: int, reads: np.ndarray, box_off: np.ndarray, idx: np.ndarray):
    """The loops of one reporter (0 D / TD, 1 SI, 2 DI, 3 INV, 4 INV_NT) -> (idx reordered per box, UniqueRead per read,
    event flags per position of the final order)"""
    reads = np.ascontiguousarray(reads, dtype=KEY_DT)
    idx = np.ascontiguousarray(idx, dtype=np.uint32).copy()
    unique = reads["unique"].copy()
    event = np.zeros(idx.size, dtype=np.uint8)
    for b in range(len(box_off) - 1):
        a, e = int(box_off[b]), int(box_off[b + 1])
        lib().bso_box_kind(reads.ctypes.data_as(C.c_void_p), unique.ctypes.data_as(C.c_void_p), C.c_void_p(idx.ctypes.data + 4 * a), C.c_uint32(e - a),
                           C.c_int(kind), C.c_void_p(event.ctypes.data + a))
    return idx, unique, event


def make_boxes_kind(kind: int, n_boxes: int, seed: int, max_reads: int = 40, heavy: int = 0):
    """Boxes for the DI / inversion loops: events with a few breakpoint pairs per box -- for the inversion kinds several pairs
    with the same BPLeft + BPRight and different sizes --, fragments seen twice (same LeftMostPos or same end, same anchor
    strand), reads of different lengths, reads already marked non-unique; (box members share every key in large groups).
    -> (reads KEY_DT, box_off, idx)"""
    rng = np.random.default_rng(seed)
    rows, box_off, idx = [], [0], []
    for b in range(n_boxes):
        n = int(rng.integers(2, max_reads + 1)) if b >= heavy else int(rng.integers(150, 400))
        base = 100_000 + 20_000 * b
        events = []
        for _ in range(int(rng.integers(1, 4))):
            bl, size = base + int(rng.integers(0, 2000)), int(rng.integers(1, 300))
            events.append((bl, bl + size, size, int(rng.integers(0, 4))))
            if kind >= 3 and rng.random() < 0.6:          # same centre, smaller span: one event for the inversion reporters
                d = int(rng.integers(1, 6))
                events.append((bl + d, bl + size - d, size - 2 * d if size > 2 * d else size, int(rng.integers(0, 4))))
        members = []
        for _ in range(n):
            bl, br, size, nt = events[int(rng.integers(0, len(events)))]
            bp = int(rng.integers(20, 26)) if rng.random() < 0.7 else int(rng.integers(5, 90))
            rl = int(rng.choice([100, 100, 100, 101, 75]))
            md = int(rng.choice([43, 45]))
            lmp = bl - bp + int(rng.integers(0, 3))
            if rng.random() < 0.2 and members:            # the same fragment again: same start (or same end), usually same strand
                src = rows[members[int(rng.integers(0, len(members)))]]
                lmp, rl = src[3], src[10]
                if rng.random() < 0.3:
                    rl2 = int(rng.choice([100, 101, 75]))
                    lmp, rl = lmp + rl - rl2, rl2          # same end, other length
                if rng.random() < 0.8:
                    md = src[9]
            row = [bl, br, size, lmp, 0, 0, nt, bp, 0 if rng.random() < 0.05 else 1, md, rl]
            members.append(len(rows))
            rows.append(tuple(row))
        order = rng.permutation(len(members))
        idx += [members[k] for k in order]
        box_off.append(len(idx))
    reads = np.zeros(len(rows), dtype=KEY_DT)
    for k, r in enumerate(rows):
        reads[k] = r
    return reads, np.array(box_off, dtype=np.uint32), np.array(idx, dtype=np.uint32)


def reference_sort_boxes(reads: np.ndarray, box_off: np.ndarray, idx: np.ndarray):
    """The same through the reference's own functions (needs oracle/_ref)."""
    from . import refhook
    L = C.CDLL(refhook.LIB)
    key = np.stack([reads[k].astype(np.int64) for k in ("bp_left", "bp_right", "indel_size", "nt_size", "bp", "left", "right", "name_id", "unique")], axis=1)
    key = (key & 0xFFFFFFFF).astype(np.uint32).view(np.int32).copy()
    idx = np.ascontiguousarray(idx, dtype=np.uint32).copy()
    box_off = np.ascontiguousarray(box_off, dtype=np.uint32)
    unique = np.zeros(reads.size, dtype=np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = L.pgref_box_sort(int(reads.size), vp(key), int(box_off.size - 1), vp(box_off), vp(idx), vp(unique))
    assert rc == 0
    return idx, unique


def make_boxes(n_boxes: int, seed: int, max_reads: int = 40, heavy: int = 0):
    """Boxes of split reads the way a window has them: a few events per box, many reads per event with equal breakpoints
    and a handful of BP values (ties everywhere), PCR duplicates (same Left/Right/Name) and both mates of a pair (same
    Name, different ends), some reads already marked non-unique.  heavy: that many boxes get hundreds of reads.
    -> (reads KEY_DT, box_off, idx)"""
    rng = np.random.default_rng(seed)
    rows, box_off, idx = [], [0], []
    name = 0
    for b in range(n_boxes):
        n = int(rng.integers(3, max_reads + 1)) if b >= heavy else int(rng.integers(200, 500))
        base = 100_000 + 20_000 * b
        events = [(base + int(rng.integers(0, 2000)), int(rng.integers(1, 300)), int(rng.integers(0, 3))) for _ in range(int(rng.integers(1, 4)))]
        members = []
        for _ in range(n):
            bl, size, nt = events[int(rng.integers(0, len(events)))]
            bp = int(rng.integers(20, 26)) if rng.random() < 0.7 else int(rng.integers(5, 90))
            left = bl - bp
            right = bl + size + 100 - bp
            kind = rng.random()
            if kind < 0.15 and members:                 # duplicate of an earlier read of this box
                src = rows[members[int(rng.integers(0, len(members)))]]
                row = list(src)
                if rng.random() < 0.5:
                    row[3] += 1                         # same name, other Left: not a duplicate
            else:
                row = [bl, bl + size, size, left, right, name, nt, bp, 1]
                name += 1
            if rng.random() < 0.05:
                row[8] = 0
            members.append(len(rows))
            rows.append(tuple(row))
        order = rng.permutation(len(members))
        idx += [members[k] for k in order]
        box_off.append(len(idx))
    reads = np.zeros(len(rows), dtype=KEY_DT)
    for k, r in enumerate(rows):
        reads[k] = (r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], 0, 0)
    return reads, np.array(box_off, dtype=np.uint32), np.array(idx, dtype=np.uint32)
