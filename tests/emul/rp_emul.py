"""tests/emul/rp_emul.py -- TEST INFRASTRUCTURE ONLY: ctypes driver of rp_emul.cpp (pg_rp.cuh compiled for the host)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from oracle import rporacle

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "librp_emul.so")
CSRC = os.path.join(HERE, "..", "..", "pindel_b200", "csrc")


def build(force=False):
    srcs = [os.path.join(HERE, "rp_emul.cpp"), os.path.join(CSRC, "pg_rp.cuh"), os.path.join(CSRC, "pg_rp_host.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-g", "-fPIC", "-shared", "-Wall", "-x", "c++", "-o", LIB, srcs[0]])
    lib = C.CDLL(LIB)
    assert lib.rpemu_sizes(0) == rporacle.RP_DT.itemsize and lib.rpemu_sizes(1) == rporacle.CLUSTER_DT.itemsize
    return lib


def cluster(lib, rows):
    rows = np.ascontiguousarray(rows, dtype=rporacle.RP_DT)
    out = np.zeros(rows.size, dtype=rporacle.CLUSTER_DT)
    lib.rpemu_modify(int(rows.size), rows.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    mod = out.copy()
    lib.rpemu_summarize(int(rows.size), rows.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return mod, out


def class_stats(lib, rows):
    """(classes of >= 2 rows, rows in them, largest class) of the product's class split."""
    rows = np.ascontiguousarray(rows, dtype=rporacle.RP_DT)
    st = np.zeros(3, dtype=np.uint32)
    lib.rpemu_class_stats(int(rows.size), rows.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p))
    return tuple(int(x) for x in st)


def cluster_interchr(lib, rows, chr_a, chr_b):
    rows = np.ascontiguousarray(rows, dtype=rporacle.RP_DT)
    ca, cb = np.ascontiguousarray(chr_a, dtype=np.int32), np.ascontiguousarray(chr_b, dtype=np.int32)
    out = np.zeros(rows.size, dtype=rporacle.CLUSTER_DT)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.rpemu_modify_interchr(int(rows.size), vp(rows), vp(ca), vp(cb), vp(out))
    mod = out.copy()
    lib.rpemu_summarize_interchr(int(rows.size), vp(rows), vp(ca), vp(cb), vp(out))
    return mod, out
