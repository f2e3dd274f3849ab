"""tests/emul/sv_emul.py -- TEST INFRASTRUCTURE ONLY: ctypes driver of sv_emul.cpp (pg_sv.cuh compiled for the host)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from pindel_b200 import runs as R, sv

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsv_emul.so")
CSRC = os.path.join(HERE, "..", "..", "pindel_b200", "csrc")


def build(force=False):
    srcs = [os.path.join(HERE, "sv_emul.cpp")] + [os.path.join(CSRC, f) for f in
                                                   ("pg_sv.cuh", "pg_sv_types.h", "pg_warp_types.h", "pg_pack.h", "pg_params.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O1", "-g", "-fPIC", "-shared", "-Wall", "-Wno-unused-function", "-x", "c++",
                               "-o", LIB, srcs[0]])
    lib = C.CDLL(LIB)
    lib.svemu_classify.restype = C.c_int64
    assert lib.svemu_record_size() == sv.RECORD_DT.itemsize
    return lib


def classify(lib, chr_seq, batch: dict, win: dict, seq_error_rate=0.01, sensitivity=0.95):
    n = len(batch["matched_d"])
    chr_seq = np.ascontiguousarray(chr_seq, dtype=np.uint8)
    reg = win.get("region")
    w = np.array([win["driver_mask"], 0, win["window_end"], win["num_boxes"], win["box_size"], 0 if reg is None else 1,
                  0 if reg is None else reg[0], 0 if reg is None else 1, 0 if reg is None else reg[1],
                  win["min_inversion_size"], win["min_matched_bases"]], dtype=np.uint32)
    arrs = [np.ascontiguousarray(batch[k]) for k in ("seq_blob", "seq_off", "matched_d", "insert_size", "rflags", "run_off",
                                                     "n_close_runs", "runs")]
    cap = 4 * n + 1024
    rec = np.zeros(cap, dtype=sv.RECORD_DT)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    k = lib.svemu_classify(vp(chr_seq), C.c_uint32(chr_seq.size), C.c_double(seq_error_rate), C.c_double(sensitivity), vp(w), n,
                           *[vp(a) for a in arrs], vp(rec), C.c_uint32(cap))
    assert 0 <= k <= cap, k
    rec = rec[:k]
    return rec[np.lexsort((rec["seq"], rec["read"], rec["driver"]))]
