"""tests/emul/sr_emul.py -- TEST INFRASTRUCTURE ONLY: ctypes driver of sr_emul.cpp (pg_sr.cuh compiled for the host)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from tests.emul.bam_emul import REC_DT

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsr_emul.so")
CSRC = os.path.join(HERE, "..", "..", "pindel_b200", "csrc")

QUERY_DT = np.dtype([("tid", "<i4"), ("beg", "<i4"), ("end", "<i4"), ("insert_size", "<i4"), ("min_anchor_quality", "<u4"), ("chr_len", "<u4"),
                     ("nm_limit", "<i4"), ("pad", "<u4"), ("max_mismatch_rate", "<f8")])
CAND_DT = np.dtype([("anchor", "<u4"), ("read", "<u4"), ("matched_rel_pos", "<u4"), ("read_length", "<u4"), ("matched_d", "u1"), ("ms", "u1"), ("pad", "<u2")])
REF_DT = np.dtype([("pos", "<u4"), ("read_length", "<i2"), ("mq", "u1"), ("pad", "u1")])


def build(force=False):
    srcs = [os.path.join(HERE, "sr_emul.cpp"), os.path.join(CSRC, "pg_sr.cuh"), os.path.join(CSRC, "pg_bam.cuh")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-g", "-fPIC", "-shared", "-Wall", "-x", "c++", "-o", LIB, srcs[0]])
    lib = C.CDLL(LIB)
    assert lib.sremu_sizes(0) == QUERY_DT.itemsize and lib.sremu_sizes(1) == CAND_DT.itemsize and lib.sremu_sizes(2) == REF_DT.itemsize
    return lib


def query(tid, beg, end, insert_size, chr_len, min_anchor_quality=0, nm_limit=2, max_mismatch_rate=0.02):
    return np.array([(tid, beg, end, insert_size, min_anchor_quality, chr_len, nm_limit, 0, max_mismatch_rate)], dtype=QUERY_DT)


def candidates(lib, stream: bytes, recs: np.ndarray, q: np.ndarray, shuffle_seed=0, table_bits=0, hash_mask=0):
    """-> (cands, seqs [n, stride] uint8, refs, records delivered, error bits)"""
    buf = np.frombuffer(stream, dtype=np.uint8)
    recs = np.ascontiguousarray(recs, dtype=REC_DT)
    counts = np.zeros(5, dtype=np.uint64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    args = (vp(buf), vp(recs), C.c_uint32(recs.size), vp(q), C.c_uint32(shuffle_seed), C.c_uint32(table_bits), C.c_uint64(hash_mask))
    rc = lib.sremu_candidates(*args, None, None, None, vp(counts))
    if rc != 0:
        raise ValueError("sremu_candidates: %d" % rc)
    n, m, stride = int(counts[0]), int(counts[1]), int(counts[2])
    cands, seqs, refs = np.zeros(n, dtype=CAND_DT), np.zeros((n, max(stride, 1)), dtype=np.uint8), np.zeros(m, dtype=REF_DT)
    if not int(counts[4]):
        lib.sremu_candidates(*args, vp(cands), vp(seqs), vp(refs), vp(counts))
    return cands, seqs, refs, int(counts[3]), int(counts[4])
