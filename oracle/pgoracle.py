"""oracle/pgoracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes wrapper around oracle/libpg_oracle.so (oracle/pg_oracle.c, the plain-C CPU restatement of the
reference's pattern-growth search).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libpg_oracle.so")


class Params(C.Structure):
    _fields_ = [("additional_mismatch", C.c_int32), ("min_close", C.c_int32), ("min_perfect_match", C.c_int32),
                ("max_range_index", C.c_int32), ("max_mismatch_rate", C.c_double), ("max_mismatch", C.c_uint32 * 500)]


def build() -> None:
    """Compile the C restatement in place (gcc only; no reference sources involved)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])


def load() -> C.CDLL:
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "pg_oracle.c")):
        build()
    lib = C.CDLL(LIB)
    lib.pgo_search_stream.restype = C.c_int64
    return lib


def default_params(lib=None, **over) -> Params:
    lib = lib or load()
    p = Params()
    lib.pgo_default_params(C.byref(p))
    seq_err = over.pop("seq_error_rate", None)
    sens = over.pop("sensitivity", None)
    if seq_err is not None or sens is not None:
        tab = (C.c_uint32 * 500)()
        lib.pgo_create_prob_table(C.c_double(0.001 + (0.01 if seq_err is None else seq_err)),
                                  C.c_double(0.95 if sens is None else sens), tab)
        p.max_mismatch = tab
    for k, v in over.items():
        setattr(p, k, v)
    return p


def prob_table(seq_error_rate=0.01, sensitivity=0.95, lib=None) -> np.ndarray:
    lib = lib or load()
    tab = np.zeros(500, dtype=np.uint32)
    lib.pgo_create_prob_table(C.c_double(0.001 + seq_error_rate), C.c_double(sensitivity), tab.ctypes.data_as(C.c_void_p))
    return tab


def search(params: Params, chr_seqs, seq_blob, seq_off, matched_d, rel_pos, insert_size, chr_id=None,
           regions=None, region_off=None, far=True, threads=1, lib=None):
    """chr_seqs: list of padded chromosome strings (uint8 arrays).  Returns (stream, (close_s, far_s));
    parse with oracle.refhook.parse_stream (same layout)."""
    lib = lib or load()
    n = len(seq_off) - 1
    chr_seqs = [np.ascontiguousarray(c, dtype=np.uint8) for c in chr_seqs]
    ptrs = (C.c_void_p * len(chr_seqs))(*[c.ctypes.data for c in chr_seqs])
    lens = np.array([c.size for c in chr_seqs], dtype=np.uint32)
    seq_blob = np.ascontiguousarray(seq_blob, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint32)
    matched_d = np.ascontiguousarray(matched_d, dtype=np.uint8)
    rel_pos = np.ascontiguousarray(rel_pos, dtype=np.uint32)
    insert_size = np.ascontiguousarray(insert_size, dtype=np.int16)
    cid = None if chr_id is None else np.ascontiguousarray(chr_id, dtype=np.int32)
    rg = None if regions is None else np.ascontiguousarray(regions, dtype=np.uint32).reshape(-1, 3)
    ro = None if region_off is None else np.ascontiguousarray(region_off, dtype=np.uint32)
    out = C.POINTER(C.c_int32)()
    secs = (C.c_double * 2)()
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    ln = lib.pgo_search_stream(C.byref(params), len(chr_seqs), ptrs, vp(lens), n, vp(seq_blob), vp(seq_off), vp(matched_d),
                               vp(rel_pos), vp(insert_size), vp(cid), vp(rg), vp(ro), 1 if far else 0, threads,
                               C.byref(out), secs)
    stream = np.ctypeslib.as_array(out, shape=(ln,)).copy()
    lib.pgo_free(out)
    return stream, (secs[0], secs[1])
