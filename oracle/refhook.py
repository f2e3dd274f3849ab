"""oracle/refhook.py -- TEST INFRASTRUCTURE ONLY.

ctypes wrapper around oracle/_ref/libpindel_ref.so: the UNMODIFIED reference objects (compiled
from /root/reference/src by oracle/Makefile) plus oracle/ref_hook.cpp.  It runs the reference's
own GetCloseEnd / CleanUniquePoints / SearchFarEnds (src/pindel.h:523,741) on a batch and returns
every UniquePoint.  The reference keeps process-global state, so use one process per configuration
(`run_in_subprocess`).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libpindel_ref.so")


def available() -> bool:
    return os.path.exists(LIB)


def write_fasta(path: str, contigs) -> None:
    """contigs: list of (name, uint8 ASCII array).  60 columns per line."""
    with open(path, "wb") as f:
        for name, seq in contigs:
            f.write(b">" + name.encode() + b"\n")
            n = seq.size
            full = (n // 60) * 60
            if full:
                body = np.empty((full // 60, 61), dtype=np.uint8)
                body[:, :60] = seq[:full].reshape(-1, 60)
                body[:, 60] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(seq[full:].tobytes() + b"\n")


class RefHook:
    """One initialised reference instance (one per process)."""

    def __init__(self, fasta: str, extra_args=(), workdir: str | None = None):
        self.lib = C.CDLL(LIB)
        L = self.lib
        L.pgref_chr_name.restype = C.c_char_p
        L.pgref_chr_seq.restype = C.c_void_p
        L.pgref_chr_len.restype = C.c_uint32
        L.pgref_setting_double.restype = C.c_double
        L.pgref_search.restype = C.c_int64
        L.pgref_classify.restype = C.c_int64
        self.workdir = workdir or tempfile.mkdtemp(prefix="pgref_")
        dummy = os.path.join(self.workdir, "empty_reads.txt")
        open(dummy, "w").close()
        args = ["-f", fasta, "-p", dummy, "-o", os.path.join(self.workdir, "out")] + [str(a) for a in extra_args]
        arr = (C.c_char_p * len(args))(*[a.encode() for a in args])
        # the reference prints progress on stdout; keep the test output clean
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)
        try:
            rc = L.pgref_init(len(args), arr)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
            os.close(devnull)
        if rc != 0:
            raise RuntimeError("pgref_init failed")
        self.n_chr = L.pgref_n_chr()

    def chr_name(self, i: int) -> str:
        return self.lib.pgref_chr_name(i).decode()

    def chr_seq(self, i: int) -> np.ndarray:
        """The padded in-memory chromosome string (src/pindel.cpp:305-309) as a uint8 array (copy)."""
        n = self.lib.pgref_chr_len(i)
        ptr = self.lib.pgref_chr_seq(i)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n,)).copy()

    def max_mismatch(self) -> np.ndarray:
        out = np.zeros(500, dtype=np.uint32)
        self.lib.pgref_max_mismatch(out.ctypes.data_as(C.c_void_p), 500)
        return out

    def setting_int(self, name: str) -> int:
        return self.lib.pgref_setting_int(name.encode())

    def setting_double(self, name: str) -> float:
        return self.lib.pgref_setting_double(name.encode())

    def update_bd(self) -> None:
        """Activate the -b BreakDancer events (BDData::UpdateBD, src/bddata.cpp:646)."""
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)
        try:
            self.lib.pgref_update_bd()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
            os.close(devnull)

    def rp_discovery(self, bam_path: str, chr_id: int, start: int, end: int, insert_size: int):
        """The reference's ReadInBamReads_RP_Discovery on one BAM file and window -> (rows [k, 9] int64: list, PosA, PosB, DA,
        DB, ReadLength, InsertSize, MQA, FNV-1a of ReadName; records scanned)."""
        self.lib.pgref_rp_discovery.restype = C.c_int64
        out = C.POINTER(C.c_int32)()
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)
        try:
            k = self.lib.pgref_rp_discovery(bam_path.encode(), chr_id, C.c_uint32(start), C.c_uint32(end), insert_size, C.byref(out))
        finally:
            os.dup2(saved, 1)
            os.close(saved)
            os.close(devnull)
        if k < 0:
            raise RuntimeError("pgref_rp_discovery failed: %d" % k)
        stream = np.ctypeslib.as_array(out, shape=(int(k),)).copy()
        self.lib.pgref_free(out)
        rows = stream[:-1].reshape(-1, 9).astype(np.int64)
        rows[:, [1, 2, 6, 8]] &= 0xFFFFFFFF
        return rows, int(stream[-1])

    def sr_candidates(self, bam_path: str, chr_id: int, start: int, end: int, insert_size: int, min_anchor_quality: int = -1):
        """What the reference's ReadInBamReads_SR hands to the ReadBuffer for one BAM file and window, before GetCloseEnd
        -> (candidates [k, 7] int64: FNV-1a of Name, MatchedD, MatchedRelPos, MS, InsertSize, ReadLength, FNV-1a of UnmatchedSeq;
        ref reads [m, 3]: Pos, MQ, ReadLength).  min_anchor_quality: the -A setting for this call (-1 = the process default)."""
        self.lib.pgref_sr_candidates.restype = C.c_int64
        out = C.POINTER(C.c_int32)()
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)
        try:
            k = self.lib.pgref_sr_candidates(bam_path.encode(), chr_id, C.c_uint32(start), C.c_uint32(end), insert_size, min_anchor_quality, C.byref(out))
        finally:
            os.dup2(saved, 1)
            os.close(saved)
            os.close(devnull)
        if k < 0:
            raise RuntimeError("pgref_sr_candidates failed: %d" % k)
        stream = np.ctypeslib.as_array(out, shape=(int(k),)).astype(np.int64)
        self.lib.pgref_free(out)
        k = int(stream[0])
        cand = stream[1:1 + 7 * k].reshape(-1, 7).copy()
        cand[:, [0, 2, 6]] &= 0xFFFFFFFF
        m = int(stream[1 + 7 * k])
        refs = stream[2 + 7 * k:2 + 7 * k + 3 * m].reshape(-1, 3).copy()
        return cand, refs

    def set_window(self, chr_id: int, start: int, end: int) -> None:
        if self.lib.pgref_set_window(chr_id, C.c_uint32(start), C.c_uint32(end)) != 0:
            raise RuntimeError("pgref_set_window failed")

    def search(self, seq_blob, seq_off, matched_d, rel_pos, insert_size, chr_id=0, far=True, threads=1, timing_only=False):
        """Returns (stream int32 array, (close_seconds, far_seconds)).  See parse_stream()."""
        n = len(seq_off) - 1
        seq_blob = np.ascontiguousarray(seq_blob, dtype=np.uint8)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint32)
        matched_d = np.ascontiguousarray(matched_d, dtype=np.uint8)
        rel_pos = np.ascontiguousarray(rel_pos, dtype=np.uint32)
        insert_size = np.ascontiguousarray(insert_size, dtype=np.int16)
        out = C.POINTER(C.c_int32)()
        secs = (C.c_double * 2)()
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)
        try:
            ln = self.lib.pgref_search(n, seq_blob.ctypes.data_as(C.c_void_p), seq_off.ctypes.data_as(C.c_void_p),
                                       matched_d.ctypes.data_as(C.c_void_p), rel_pos.ctypes.data_as(C.c_void_p),
                                       insert_size.ctypes.data_as(C.c_void_p), chr_id, (3 if far else 1) | (4 if timing_only else 0), threads,
                                       C.byref(out), secs)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
            os.close(devnull)
        if ln < 0:
            raise RuntimeError("pgref_search failed")
        stream = np.ctypeslib.as_array(out, shape=(ln,)).copy()
        self.lib.pgref_free(out)
        return stream, (secs[0], secs[1])


    def classify(self, seq_blob, seq_off, matched_d, insert_size, rflags, up_rows, up_off, n_close, window, num_boxes, box_size,
                 chr_id=0, driver_mask=0x7F, region=None):
        """The reference's SV drivers on one window.  up_rows int32 [k,5] (len, absLoc, dir, strand, mm); window = (start, end).
        Returns (dump stream int32, seconds[7]); parse with oracle.svoracle.parse_dump."""
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
        secs = (C.c_double * 7)()
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)
        try:
            vp = lambda a: a.ctypes.data_as(C.c_void_p)
            ln = self.lib.pgref_classify(n, vp(seq_blob), vp(seq_off), vp(matched_d), vp(insert_size), vp(rflags), vp(up_rows),
                                         vp(up_off), vp(n_close), chr_id, C.c_uint32(window[0]), C.c_uint32(window[1]),
                                         C.c_uint32(num_boxes), C.c_uint32(box_size), C.c_uint32(driver_mask),
                                         0 if region is None else 1, C.c_uint32(0 if region is None else region[0]),
                                         C.c_uint32(0 if region is None else region[1]), C.byref(out), secs)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
            os.close(devnull)
        if ln < 0:
            raise RuntimeError("pgref_classify failed")
        stream = np.ctypeslib.as_array(out, shape=(max(int(ln), 1),))[:ln].copy()
        self.lib.pgref_free(out)
        return stream, [secs[i] for i in range(7)]


def parse_stream(stream: np.ndarray, n: int):
    """-> list of dicts {flip, L, close_raw, close, far, regions}; UPs as int arrays [k,6]
    (len, absLoc, dir 0=FORWARD/1=BACKWARD, strand 0=SENSE/1=ANTISENSE, mm, chr)."""
    out = []
    p = 0
    s = stream
    for _ in range(n):
        flip, L, ncr, ncc, nf, nreg = (int(x) for x in s[p:p + 6])
        p += 6
        cr = s[p:p + 6 * ncr].reshape(ncr, 6).astype(np.int64); p += 6 * ncr
        cc = s[p:p + 6 * ncc].reshape(ncc, 6).astype(np.int64); p += 6 * ncc
        fa = s[p:p + 6 * nf].reshape(nf, 6).astype(np.int64); p += 6 * nf
        rg = s[p:p + 3 * nreg].reshape(nreg, 3).astype(np.int64); p += 3 * nreg
        for a in (cr, cc, fa):
            a[:, 1] &= 0xFFFFFFFF  # AbsLoc is unsigned in the reference
        out.append(dict(flip=flip, L=L, close_raw=cr, close=cc, far=fa, regions=rg))
    assert p == s.size
    return out


def rp_cluster(rows):
    """The reference's own sort + ModifyRP + Summarize (oracle/ref_hook.cpp pgref_rp_cluster; needs no initialised
    instance).  rows: structured array with pos_a, pos_b, da, db, read_length, insert_size (oracle.rporacle.RP_DT).
    -> (table [n, 14] int64: input index, OriginalPosA, OriginalPosB, DA, DB, ReadLength, InsertSize, PosA, PosA1, PosB, PosB1,
    Visited, Report, NumberOfIdentical -- in the order ModifyRP's sort leaves; tags: list of input-index lists; seconds)."""
    lib = C.CDLL(LIB)
    lib.pgref_rp_cluster.restype = C.c_int64
    n = len(rows)
    cols = [np.ascontiguousarray(rows[k]) for k in ("pos_a", "pos_b", "da", "db", "read_length", "insert_size")]
    out = C.POINTER(C.c_int32)()
    secs = (C.c_double * 2)()
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)
    try:
        k = lib.pgref_rp_cluster(n, *[c.ctypes.data_as(C.c_void_p) for c in cols], C.byref(out), secs)
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)
    stream = np.ctypeslib.as_array(out, shape=(max(int(k), 1),))[:k].copy()
    lib.pgref_free(out)
    table = stream[:14 * n].reshape(n, 14).astype(np.int64)
    table[:, [1, 2, 6, 7, 8, 9, 10, 13]] &= 0xFFFFFFFF
    tags, p = [], 14 * n
    for _ in range(n):
        c = int(stream[p])
        tags.append(stream[p + 1:p + 1 + c].tolist())
        p += 1 + c
    return table, tags, (secs[0], secs[1])


def rp_cluster_interchr(rows, chr_a, chr_b):
    """The reference's own ModifyRP_InterChr + Summarize_InterChr (oracle/ref_hook.cpp pgref_rp_cluster_interchr).
    -> (table [n, 14] as rp_cluster, tags, chr_a, chr_b in the order ModifyRP_InterChr's sort leaves)"""
    lib = C.CDLL(LIB)
    lib.pgref_rp_cluster_interchr.restype = C.c_int64
    n = len(rows)
    cols = [np.ascontiguousarray(rows[k]) for k in ("pos_a", "pos_b", "da", "db", "read_length", "insert_size")]
    cols += [np.ascontiguousarray(chr_a, dtype=np.int32), np.ascontiguousarray(chr_b, dtype=np.int32)]
    out = C.POINTER(C.c_int32)()
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)
    try:
        k = lib.pgref_rp_cluster_interchr(n, *[c.ctypes.data_as(C.c_void_p) for c in cols], C.byref(out))
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)
    stream = np.ctypeslib.as_array(out, shape=(max(int(k), 1),))[:k].copy()
    lib.pgref_free(out)
    table = stream[:14 * n].reshape(n, 14).astype(np.int64)
    table[:, [1, 2, 6, 7, 8, 9, 10, 13]] &= 0xFFFFFFFF
    tags, p = [], 14 * n
    for _ in range(n):
        c = int(stream[p])
        tags.append(stream[p + 1:p + 1 + c].tolist())
        p += 1 + c
    chrs = stream[p:p + 2 * n].reshape(n, 2)
    return table, tags, chrs[:, 0].copy(), chrs[:, 1].copy()
