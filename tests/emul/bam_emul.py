"""tests/emul/bam_emul.py -- TEST INFRASTRUCTURE ONLY: ctypes driver of bam_emul.cpp (pg_bam.cuh compiled for the host)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libbam_emul.so")
CSRC = os.path.join(HERE, "..", "..", "pindel_b200", "csrc")

REC_DT = np.dtype([("off", "<u8"), ("tid", "<i4"), ("pos", "<i4"), ("end", "<i4"), ("flag", "<u2"), ("mapq", "u1"), ("pad", "u1")])
assert REC_DT.itemsize == 24


def build(force=False):
    srcs = [os.path.join(HERE, "bam_emul.cpp"), os.path.join(CSRC, "pg_bam.cuh"), os.path.join(CSRC, "pg_bam_host.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-g", "-fPIC", "-shared", "-Wall", "-x", "c++", "-o", LIB, srcs[0]])
    lib = C.CDLL(LIB)
    lib.bamemu_index.restype = C.c_long
    lib.bamemu_index.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.bamemu_index_run.restype = C.c_long
    lib.bamemu_index_run.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int32, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.bamemu_last_walks.restype = C.c_uint64
    assert lib.bamemu_rec_size() == REC_DT.itemsize
    return lib


def index(lib, stream: bytes, seg: int = 0):
    """-> (records, stats (segments, segments walked again, first record offset))"""
    buf = np.frombuffer(stream, dtype=np.uint8)
    stats = np.zeros(3, dtype=np.uint64)
    n = lib.bamemu_index(buf.ctypes.data, buf.size, seg, None, 0, stats.ctypes.data)
    if n < 0:
        raise ValueError("bamemu_index: %d" % n)
    recs = np.zeros(n, dtype=REC_DT)
    n2 = lib.bamemu_index(buf.ctypes.data, buf.size, seg, recs.ctypes.data, n, stats.ctypes.data)
    assert n2 == n
    return recs, tuple(int(x) for x in stats)


def index_run(lib, run: bytes, first_record: int, n_ref: int, seg: int = 0):
    """a headerless run of blocks -> (records, tail offset, segments walked again)"""
    buf = np.frombuffer(run, dtype=np.uint8)
    stats = np.zeros(4, dtype=np.uint64)
    n = lib.bamemu_index_run(buf.ctypes.data, buf.size, first_record, n_ref, seg, None, 0, stats.ctypes.data)
    if n < 0:
        raise ValueError("bamemu_index_run: %d" % n)
    recs = np.zeros(n, dtype=REC_DT)
    assert lib.bamemu_index_run(buf.ctypes.data, buf.size, first_record, n_ref, seg, recs.ctypes.data, n, stats.ctypes.data) == n
    return recs, int(stats[3]), int(stats[1])


RP_DT = np.dtype([("pos_a", "<u4"), ("pos_b", "<u4"), ("insert_size", "<u4"), ("read_length", "<i2"), ("da", "u1"), ("db", "u1")])
QUERY_DT = np.dtype([("tid", "<i4"), ("beg", "<i4"), ("end", "<i4"), ("insert_size", "<i4"), ("min_anchor_quality", "<u4")])


def rp_discovery(lib, stream: bytes, recs: np.ndarray, tid, beg, end, insert_size, min_anchor_quality=0):
    """-> (rows, src, inter_rows, inter_src, records delivered)"""
    buf = np.frombuffer(stream, dtype=np.uint8)
    recs = np.ascontiguousarray(recs, dtype=REC_DT)
    q = np.array([(tid, beg, end, insert_size, min_anchor_quality)], dtype=QUERY_DT)
    n = recs.size
    rows, src = np.zeros(n, dtype=RP_DT), np.zeros(n, dtype=np.uint32)
    irows, isrc = np.zeros(n, dtype=RP_DT), np.zeros(n, dtype=np.uint32)
    counts = np.zeros(3, dtype=np.uint64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.bamemu_rp_discovery(vp(buf), vp(recs), C.c_uint32(n), vp(q), vp(rows), vp(src), vp(irows), vp(isrc), vp(counts))
    a, b = int(counts[0]), int(counts[1])
    return rows[:a], src[:a], irows[:b], isrc[:b], int(counts[2])
