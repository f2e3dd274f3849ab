"""tests/emul/emul.py -- TEST INFRASTRUCTURE ONLY: ctypes driver for the CPU lane emulator (pg_emul.cpp),
which executes the CUDA warp program of pindel_b200/csrc/pg_warp.cuh on the host, one warp at a time."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from pindel_b200 import runs as R

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libpg_emul.so")
CSRC = os.path.join(HERE, "..", "..", "pindel_b200", "csrc")


def build(force=False):
    srcs = [os.path.join(HERE, "pg_emul.cpp")] + [os.path.join(CSRC, f) for f in
                                                   ("pg_warp.cuh", "pg_warp_types.h", "pg_pack.h", "pg_params.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-g", "-fPIC", "-shared", "-pthread", "-Wall", "-Wno-unused-function", "-Wno-maybe-uninitialized",
                               "-Wno-unknown-pragmas"] + os.environ.get("PG_EMUL_CXXFLAGS", "").split() + ["-o", LIB, srcs[0]])
    return C.CDLL(LIB)


def make_params(lib, seq_error_rate=0.01, sensitivity=0.95, max_mismatch_rate=0.02, add_mm=1, min_close=8,
                min_perfect=3, max_range_index=2):
    blob = np.zeros(lib.pgemu_sizes(0), dtype=np.uint8)
    rc = lib.pgemu_make_params(blob.ctypes.data_as(C.c_void_p), C.c_double(seq_error_rate), C.c_double(sensitivity),
                               C.c_double(max_mismatch_rate), add_mm, min_close, min_perfect, max_range_index)
    assert rc == 0
    return blob


def search(lib, params, chr_seqs, seq_blob, seq_off, matched_d, rel_pos, insert_size, chr_id=None, mode=3,
           close_info=None, regions=None, region_off=None, pool_per_read=64):
    assert lib.pgemu_sizes(1) == R.RUN_DT.itemsize and lib.pgemu_sizes(2) == R.OUT_DT.itemsize
    n = len(seq_off) - 1
    chr_seqs = [np.ascontiguousarray(c, dtype=np.uint8) for c in chr_seqs]
    ptrs = (C.c_void_p * len(chr_seqs))(*[c.ctypes.data for c in chr_seqs])
    lens = np.array([c.size for c in chr_seqs], dtype=np.uint32)
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    seq_blob = np.ascontiguousarray(seq_blob, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint32)
    matched_d = np.ascontiguousarray(matched_d, dtype=np.uint8)
    rel_pos = np.ascontiguousarray(rel_pos, dtype=np.uint32)
    insert_size = np.ascontiguousarray(insert_size, dtype=np.int16)
    cid = None if chr_id is None else np.ascontiguousarray(chr_id, dtype=np.int32)
    ci = None if close_info is None else np.ascontiguousarray(close_info, dtype=R.CLOSEINFO_DT)
    rg = None if regions is None else np.ascontiguousarray(regions, dtype=R.REGION_DT)
    ro = None if region_off is None else np.ascontiguousarray(region_off, dtype=np.uint32)
    out = np.zeros(n, dtype=R.OUT_DT)
    cap = max(1024, n * pool_per_read)
    pool = np.zeros(cap, dtype=R.RUN_DT)
    used = C.c_uint32(0)
    rc = lib.pgemu_search(len(chr_seqs), ptrs, vp(lens), vp(params), n, mode, vp(seq_blob), vp(seq_off), vp(matched_d),
                          vp(rel_pos), vp(insert_size), vp(cid), vp(ci), vp(rg), vp(ro), vp(out), vp(pool), cap,
                          C.byref(used))
    assert rc == 0, "run pool overflow"
    return out, pool[:used.value]


def filter_stats(lib):
    """(seeds that entered the far rings' threshold filter, survivors, tiles filtered) since the last call."""
    a = (C.c_ulonglong * 3)()
    lib.pgemu_filter_stats(a)
    return int(a[0]), int(a[1]), int(a[2])
