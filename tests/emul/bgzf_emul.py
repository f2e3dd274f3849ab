"""tests/emul/bgzf_emul.py -- TEST INFRASTRUCTURE ONLY: ctypes driver of bgzf_emul.cpp (pg_bgzf.cuh compiled for the host)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libbgzf_emul.so")
CSRC = os.path.join(HERE, "..", "..", "pindel_b200", "csrc")


def build(force=False):
    srcs = [os.path.join(HERE, "bgzf_emul.cpp"), os.path.join(CSRC, "pg_bgzf.cuh"), os.path.join(CSRC, "pg_bgzf_host.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-g", "-fPIC", "-shared", "-Wall", "-x", "c++", "-o", LIB, srcs[0]])
    lib = C.CDLL(LIB)
    lib.bgzfemu_size.restype = C.c_long
    lib.bgzfemu_size.argtypes = [C.c_void_p, C.c_size_t]
    lib.bgzfemu_inflate.restype = C.c_long
    lib.bgzfemu_inflate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_long, C.c_int]
    lib.bgzfemu_inflate_raw.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
    return lib


def inflate(lib, data: bytes, check_crc: bool = True):
    """-> (inflated bytes, status per block)"""
    buf = np.frombuffer(data + b"\0" * 16, dtype=np.uint8).copy()
    total = lib.bgzfemu_size(buf.ctypes.data, len(data))
    if total < 0:
        raise ValueError("not BGZF")
    out = np.zeros(total + 16, dtype=np.uint8)
    status = np.full(1 << 20, 99, dtype=np.uint32)
    nb = lib.bgzfemu_inflate(buf.ctypes.data, len(data), out.ctypes.data, status.ctypes.data, status.size, int(check_crc))
    return out[:total].tobytes(), status[:nb]


def inflate_raw(lib, comp: bytes, isize: int):
    buf = np.frombuffer(b"\0" * 4 + comp + b"\0" * 16, dtype=np.uint8).copy()
    out = np.zeros(isize + 16, dtype=np.uint8)
    produced = C.c_uint32(0)
    rc = lib.bgzfemu_inflate_raw(buf.ctypes.data + 4, len(comp), out.ctypes.data, isize, C.byref(produced))
    return rc, out[:produced.value].tobytes()
