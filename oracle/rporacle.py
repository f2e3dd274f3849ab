"""ctypes wrapper of oracle/rp_oracle.c (TEST INFRASTRUCTURE ONLY): the CPU restatement of ModifyRP / Summarize
(src/bddata.cpp:385-431, 470-575) on rows that are already in the order ModifyRP's sort leaves."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

RP_DT = np.dtype([("pos_a", "<u4"), ("pos_b", "<u4"), ("insert_size", "<u4"), ("read_length", "<i2"), ("da", "u1"), ("db", "u1")])
CLUSTER_DT = np.dtype([("pos_a", "<u4"), ("pos_a1", "<u4"), ("pos_b", "<u4"), ("pos_b1", "<u4"), ("n_identical", "<u4"),
                       ("merged_into", "<u4"), ("visited", "u1"), ("report", "u1"), ("pad", "<u2")])
assert RP_DT.itemsize == 16 and CLUSTER_DT.itemsize == 28


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "librp_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/librp_oracle.so missing: run `make -C oracle port`")
        _lib = C.CDLL(path)
    return _lib


def modify(rows: np.ndarray) -> np.ndarray:
    rows = np.ascontiguousarray(rows, dtype=RP_DT)
    out = np.zeros(rows.size, dtype=CLUSTER_DT)
    lib().rpo_modify(int(rows.size), rows.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def summarize(rows: np.ndarray, state: np.ndarray) -> np.ndarray:
    rows = np.ascontiguousarray(rows, dtype=RP_DT)
    out = np.ascontiguousarray(state, dtype=CLUSTER_DT).copy()
    lib().rpo_summarize(int(rows.size), rows.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def cluster(rows: np.ndarray) -> np.ndarray:
    return summarize(rows, modify(rows))


def cluster_interchr(rows: np.ndarray, chr_a: np.ndarray, chr_b: np.ndarray):
    """ModifyRP_InterChr (after its sort) + Summarize_InterChr, src/bddata.cpp:439-468, 577-624.  -> (after modify, after summarize)"""
    rows = np.ascontiguousarray(rows, dtype=RP_DT)
    ca, cb = np.ascontiguousarray(chr_a, dtype=np.int32), np.ascontiguousarray(chr_b, dtype=np.int32)
    out = np.zeros(rows.size, dtype=CLUSTER_DT)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().rpo_modify_interchr(int(rows.size), vp(rows), vp(ca), vp(cb), vp(out))
    mod = out.copy()
    lib().rpo_summarize_interchr(int(rows.size), vp(rows), vp(ca), vp(cb), vp(out))
    return mod, out
