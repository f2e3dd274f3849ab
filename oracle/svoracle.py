"""ctypes wrapper of oracle/sv_oracle.c (TEST INFRASTRUCTURE ONLY) + the parser of the driver dumps that both the C
restatement (svo_classify) and the reference hook (pgref_classify, oracle/ref_hook.cpp) produce."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
DRIVERS = ("D", "DI", "TD", "TD_NT", "INV", "INV_NT", "SI")
STATE_KEYS = ("Left", "Right", "BP", "BPLeft", "BPRight", "IndelSize", "NT_size", "Used", "nt_len", "nt_hash", "nt_head", "nt_tail")


class SvParams(C.Structure):
    _fields_ = [("driver_mask", C.c_uint32), ("window_end", C.c_uint32), ("num_boxes", C.c_uint32), ("box_size", C.c_uint32),
                ("region_start_defined", C.c_uint32), ("region_start", C.c_uint32), ("region_end_defined", C.c_uint32),
                ("region_end", C.c_uint32), ("min_inv", C.c_int32), ("min_matched", C.c_int32), ("seq_error_rate", C.c_double),
                ("max_snp", C.c_int32 * 1024)]


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libsv_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/libsv_oracle.so missing: run `make -C oracle port`")
        _lib = C.CDLL(path)
        _lib.svo_classify.restype = C.c_int64
        _lib.svo_classify.argtypes = [C.POINTER(SvParams), C.c_char_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(C.c_int32))]
        _lib.svo_free.argtypes = [C.c_void_p]
    return _lib


def make_params(max_snp, window_end, num_boxes, box_size, driver_mask=0x7F, region=None, min_inv=30, min_matched=30,
                seq_error_rate=0.01):
    P = SvParams()
    P.driver_mask, P.window_end, P.num_boxes, P.box_size = driver_mask, window_end, num_boxes, box_size
    if region is not None:
        (sd, s), (ed, e) = region
        P.region_start_defined, P.region_start, P.region_end_defined, P.region_end = int(sd), s, int(ed), e
    P.min_inv, P.min_matched, P.seq_error_rate = min_inv, min_matched, seq_error_rate
    for i, v in enumerate(max_snp[:1024]):
        P.max_snp[i] = int(v)
    return P


def classify(P, chr_seq: bytes, seq_blob, seq_off, matched_d, insert_size, rflags, up_rows, up_off, n_close):
    """up_rows: int32 [k, 5] (len, absLoc, dir, strand, mm).  Returns the parsed dump (see parse_dump)."""
    n = len(matched_d)
    seq_blob = np.ascontiguousarray(seq_blob, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint32)
    matched_d = np.ascontiguousarray(matched_d, dtype=np.uint8)
    insert_size = np.ascontiguousarray(insert_size, dtype=np.int16)
    rflags = np.ascontiguousarray(rflags, dtype=np.uint8)
    up_rows = np.ascontiguousarray(up_rows, dtype=np.int32).reshape(-1, 5)
    up_off = np.ascontiguousarray(up_off, dtype=np.uint32)
    n_close = np.ascontiguousarray(n_close, dtype=np.uint32)
    out = C.POINTER(C.c_int32)()
    k = lib().svo_classify(C.byref(P), chr_seq, len(chr_seq), n, seq_blob.ctypes.data, seq_off.ctypes.data, matched_d.ctypes.data,
                           insert_size.ctypes.data, rflags.ctypes.data, up_rows.ctypes.data, up_off.ctypes.data,
                           n_close.ctypes.data, C.byref(out))
    stream = np.ctypeslib.as_array(out, shape=(max(int(k), 1),))[:k].copy()
    lib().svo_free(out)
    return parse_dump(stream, n)


def parse_dump(stream: np.ndarray, n: int):
    """-> {driver: dict(vars=[k,2] (box, read), future=[m,13] (read + state), state=[n,12])}"""
    res, p = {}, 0
    while p < len(stream):
        d, nv, nf = int(stream[p]), int(stream[p + 1]), int(stream[p + 2])
        p += 3
        v = stream[p:p + 2 * nv].reshape(nv, 2)
        p += 2 * nv
        f = stream[p:p + 13 * nf].reshape(nf, 13)
        p += 13 * nf
        s = stream[p:p + 12 * n].reshape(n, 12)
        p += 12 * n
        res[d] = dict(vars=v, future=f, state=s)
    return res


def compare_dumps(a, b, what=""):
    assert sorted(a) == sorted(b), (what, sorted(a), sorted(b))
    for d in sorted(a):
        for key in ("vars", "future", "state"):
            x, y = a[d][key], b[d][key]
            if x.shape != y.shape or not np.array_equal(x, y):
                msg = "%s: driver %s: %s differ (%s vs %s)" % (what, DRIVERS[d], key, x.shape, y.shape)
                if x.shape == y.shape:
                    bad = np.nonzero((x != y).any(axis=1))[0]
                    msg += "; %d rows, first row %d:\n  %s\n  %s" % (bad.size, bad[0], x[bad[0]].tolist(), y[bad[0]].tolist())
                else:
                    m = min(len(x), len(y))
                    bad = np.nonzero((x[:m] != y[:m]).any(axis=1))[0]
                    if bad.size:
                        msg += "; first differing row %d:\n  %s\n  %s" % (bad[0], x[bad[0]].tolist(), y[bad[0]].tolist())
                raise AssertionError(msg)
