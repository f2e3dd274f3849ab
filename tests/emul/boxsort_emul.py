"""tests/emul/boxsort_emul.py -- TEST INFRASTRUCTURE ONLY: ctypes driver of boxsort_emul.cpp (pg_boxsort.cuh for the host)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from oracle import boxsortoracle as B

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libboxsort_emul.so")
CSRC = os.path.join(HERE, "..", "..", "pindel_b200", "csrc")


def build(force=False):
    srcs = [os.path.join(HERE, "boxsort_emul.cpp"), os.path.join(CSRC, "pg_boxsort.cuh")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-g", "-fPIC", "-shared", "-Wall", "-x", "c++", "-o", LIB, srcs[0]])
    lib = C.CDLL(LIB)
    assert lib.boxemu_key_size() == B.KEY_DT.itemsize
    return lib


def sort_boxes(lib, member_keys: np.ndarray, box_off: np.ndarray, warp_formulation: bool = False):
    """member_keys: KEY_DT per member (already gathered in box order).  -> (order within each box, nonunique per sorted position)"""
    keys = np.ascontiguousarray(member_keys, dtype=B.KEY_DT)
    box_off = np.ascontiguousarray(box_off, dtype=np.uint32)
    order = np.zeros(keys.size, dtype=np.uint32)
    non = np.zeros(keys.size, dtype=np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    fn = lib.boxemu_sort_warp_formulation if warp_formulation else lib.boxemu_sort
    fn(C.c_uint32(box_off.size - 1), vp(box_off), vp(keys), vp(order), vp(non))
    return order, non


def sort_boxes_kind(lib, kind: int, member_keys: np.ndarray, box_off: np.ndarray, warp_formulation: bool = False):
    """-> (order within each box, nonunique per sorted position, event flags per sorted position)"""
    keys = np.ascontiguousarray(member_keys, dtype=B.KEY_DT)
    box_off = np.ascontiguousarray(box_off, dtype=np.uint32)
    order = np.zeros(keys.size, dtype=np.uint32)
    non = np.zeros(keys.size, dtype=np.uint8)
    ev = np.zeros(keys.size, dtype=np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    fn = lib.boxemu_sort_warp_formulation_kind if warp_formulation else lib.boxemu_sort_kind
    fn(C.c_int(kind), C.c_uint32(box_off.size - 1), vp(box_off), vp(keys), vp(order), vp(non), vp(ev))
    return order, non, ev
