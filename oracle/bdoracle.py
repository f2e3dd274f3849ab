"""oracle/bdoracle.py -- TEST INFRASTRUCTURE ONLY.

Python restatement of the reference's BreakDancer mask and cluster lookup -- BDData::loadRegion (src/bddata.cpp:850-946),
BDData::createRegionCluster (:820-846), BDData::getCorrespondingSearchWindowCluster (:949-967), with the coordinate
helpers of src/control_state.cpp:25-132 -- written position by position like the reference (slow; for test-sized windows).
Pinned against the reference's own functions through oracle/_ref/libpindel_ref.so (pgref_bd_clusters, which drives them on
a caller-supplied event list) by tests/test_bd_oracle_pinning.py.  Only tests/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import functools

import numpy as np

SPAN = 200  # BreakDancerCoordinate::BREAKDANCER_WINDOWSPAN, src/control_state.h:46


def normalise(events):
    """BreakDancerCoordinate(name, pos, pos2) orders its two positions (src/control_state.cpp:25-34)."""
    ev = np.array(events, dtype=np.int64).reshape(-1, 6).copy()
    for a in (1, 4):
        lo = np.minimum(ev[:, a], ev[:, a + 1])
        hi = np.maximum(ev[:, a], ev[:, a + 1])
        ev[:, a], ev[:, a + 1] = lo, hi
    return ev


def start_of_window(pos, pos2):      # src/control_state.cpp:71-83 (position2 < position never holds after the constructor)
    return pos - SPAN if pos >= SPAN else 0


def end_of_window(pos, pos2):        # src/control_state.cpp:90-98
    return (pos2 if (pos2 > pos and pos2 > 0) else pos) + SPAN


def _cmp_coord(names, ca, pa, cb, pb):   # operator< / operator!=, src/control_state.cpp:100-109
    if names[ca] != names[cb]:
        return -1 if names[ca] < names[cb] else 1
    if pa != pb:
        return -1 if pa < pb else 1
    return 0


def sort_on_first(ev, names):        # sortOnFirstBDCoordinate, src/control_state.cpp:111-120
    def cmp(x, y):
        c = _cmp_coord(names, x[0], x[1], y[0], y[1])
        return c if c else _cmp_coord(names, x[3], x[4], y[3], y[4])
    return sorted([tuple(int(v) for v in r) for r in ev], key=functools.cmp_to_key(cmp))


def sort_on_second(rows, names):     # sortOnSecondBDCoordinate, src/control_state.cpp:122-131
    def cmp(x, y):
        c = _cmp_coord(names, x[3], x[4], y[3], y[4])
        return c if c else _cmp_coord(names, x[0], x[1], y[0], y[1])
    return sorted(rows, key=functools.cmp_to_key(cmp))


def create_region_cluster(rows, names):
    """src/bddata.cpp:820-846 -> [(chromosome index, start, end)]"""
    sub = sort_on_second(rows, names)
    out = []
    i = 0
    while i < len(sub):
        e = sub[i]
        cur = [e[3], start_of_window(e[4], e[5]), end_of_window(e[4], e[5])]
        while i + 1 < len(sub) and names[sub[i + 1][3]] == names[sub[i][3]] and \
                start_of_window(sub[i + 1][4], sub[i + 1][5]) <= end_of_window(sub[i][4], sub[i][5]) + 1:     # regionsOverlap, :814-818
            i += 1
            cur[2] = end_of_window(sub[i][4], sub[i][5])
        out.append(tuple(cur))
        i += 1
    return out


def clusters(events, names, chr_id, start, end, queries):
    """loadRegion for SearchWindow(chr_id, start, end), then getCorrespondingSearchWindowCluster for every close-end
    position in `queries` -> list of [(chromosome index, start, end)] (None where the reference reads past its mask)."""
    ev = sort_on_first(normalise(events), names)
    INS = 1000
    ws = start - 3 * INS if start >= 3 * INS else 0                              # :857-863
    we = end + 3 * INS                                                           # :872

    # lower_bound / upper_bound with (chromosome, position) and an empty second coordinate (:879-885)
    def less_than_probe(e, pos):       # e < probe
        c = _cmp_coord(names, e[0], e[1], chr_id, pos)
        return c < 0 if c else False   # equal firsts: e.second < ("", 0) is never true

    def probe_less_than(e, pos):       # probe < e
        c = _cmp_coord(names, chr_id, pos, e[0], e[1])
        return c < 0 if c else True    # equal firsts: ("", 0) < e.second always

    lo = next((i for i, e in enumerate(ev) if not less_than_probe(e, ws)), len(ev))
    hi = next((i for i, e in enumerate(ev) if probe_less_than(e, we)), len(ev))
    size = 1 + we - ws                 # SearchWindow::getSize, src/pindel.h:696-699; the mask has `size` entries (:896) ...
    mask = np.zeros(size, dtype=np.int64)
    collection = [[]]
    s_idx = e_idx = lo
    index = 0
    for p in range(ws, we):                                                      # :904-944 ... of which [0, size - 1) are written
        changed = False
        it = s_idx
        while it < e_idx:
            if p > end_of_window(ev[it][1], ev[it][2]):
                s_idx += 1
                changed = True
            else:
                break
            it += 1
        it = e_idx
        while it < hi:
            st = start_of_window(ev[it][1], ev[it][2])
            if p < st:
                break
            elif p == st:
                e_idx += 1
                changed = True
            it += 1
        if s_idx == e_idx:
            mask[p - ws] = 0
        else:
            if changed:
                index += 1
                collection.append(create_region_cluster(ev[s_idx:e_idx], names))
            mask[p - ws] = index
    out = []
    for q in queries:                                                            # :949-967 (unsigned arithmetic)
        q = int(q)
        d = (q - ws) & 0xFFFFFFFF
        if d > size:
            out.append([])
            continue
        ci = 0
        if q > ws:
            if d >= size - 1:
                out.append(None)      # reads the mask's last entry, which the sweep never writes, or one past the array: undefined
                continue
            ci = int(mask[d])
        out.append(list(collection[ci]))
    return out


def reference_clusters(hook, events, chr_id, start, end, queries):
    """The reference's own functions through the hook library (hook: an initialised oracle.refhook.RefHook)."""
    lib = hook.lib
    lib.pgref_bd_clusters.restype = C.c_int64
    ev = np.ascontiguousarray(np.array(events, dtype=np.int32).reshape(-1, 6))
    q = np.ascontiguousarray(queries, dtype=np.uint32)
    out = C.POINTER(C.c_int32)()
    n = lib.pgref_bd_clusters(ev.shape[0], ev.ctypes.data_as(C.c_void_p), chr_id, C.c_uint32(start), C.c_uint32(end), q.size,
                              q.ctypes.data_as(C.c_void_p), C.byref(out))
    s = np.ctypeslib.as_array(out, shape=(max(int(n), 1),))[:n].copy()
    lib.pgref_free(out)
    res, p = [], 0
    for _ in range(q.size):
        c = int(s[p])
        p += 1
        res.append([(int(s[p + 3 * j]), int(s[p + 3 * j + 1]) & 0xFFFFFFFF, int(s[p + 3 * j + 2]) & 0xFFFFFFFF) for j in range(c)])
        p += 3 * c
    return res
