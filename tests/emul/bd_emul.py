"""tests/emul/bd_emul.py -- TEST INFRASTRUCTURE ONLY: ctypes driver of the host build of pg_bd.cuh / pg_bd_host.h."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libbd_emul.so")
CSRC = os.path.join(HERE, "..", "..", "pindel_b200", "csrc")


def build(force=False):
    srcs = [os.path.join(HERE, "bd_emul.cpp")] + [os.path.join(CSRC, f) for f in ("pg_bd.cuh", "pg_bd_host.h", "pg_warp_types.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O1", "-g", "-fPIC", "-shared", "-Wall", "-x", "c++", "-o", LIB, srcs[0]])
    return C.CDLL(LIB)


def name_ranks(names):
    order = sorted(range(len(names)), key=lambda i: names[i])
    r = np.zeros(len(names), dtype=np.int32)
    for k, i in enumerate(order):
        r[i] = k
    return r


def clusters(lib, events_sorted, names, chr_id, start, end, queries):
    """events_sorted: [n, 6] in the reference's order -> list of [(chr, start, end)] per query"""
    ev = np.ascontiguousarray(np.array(events_sorted, dtype=np.int32).reshape(-1, 6))
    q = np.ascontiguousarray(queries, dtype=np.uint32)
    ranks = name_ranks(names)
    off = np.zeros(q.size + 1, dtype=np.uint32)
    cap = 64 * q.size + 64
    reg = np.zeros((cap, 3), dtype=np.int32)
    lib.bdemu_clusters.restype = C.c_long
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    t = lib.bdemu_clusters(vp(ev), ev.shape[0], vp(ranks), len(names), chr_id, C.c_uint32(start), C.c_uint32(end), q.size, vp(q), vp(off), vp(reg), cap)
    assert 0 <= t <= cap, t
    return [[(int(c), int(s) & 0xFFFFFFFF, int(e) & 0xFFFFFFFF) for c, s, e in reg[off[i]:off[i + 1]]] for i in range(q.size)]
