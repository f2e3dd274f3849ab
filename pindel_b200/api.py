"""Host-side mirror of the reference's search seam on top of libpindel_pg.so (ctypes over the C ABI).

Names follow the reference: `get_close_end` is the batched GetCloseEnd (src/pindel.h:523), `search_far_ends`
the batched SearchFarEnds (src/pindel.h:741), `search` runs both on the device.  This module contains no
search logic: if the CUDA library is missing or no device is present it raises -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import runs as R

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PG_LIB", os.path.join(HERE, "libpindel_pg.so"))

PG_OK, PG_ERR_ARG, PG_ERR_CUDA, PG_ERR_STATE, PG_ERR_RUNS_CAP, PG_ERR_UNSUPPORTED = range(6)


class Settings(C.Structure):
    _fields_ = [("seq_error_rate", C.c_double), ("sensitivity", C.c_double), ("max_mismatch_rate", C.c_double),
                ("additional_mismatch", C.c_int32), ("min_close", C.c_int32), ("min_perfect_match", C.c_int32),
                ("max_range_index", C.c_int32), ("device", C.c_int32), ("reserved", C.c_int32)]


class _Reads(C.Structure):
    _fields_ = [("n", C.c_int32), ("seq_blob", C.c_void_p), ("seq_off", C.c_void_p), ("matched_d", C.c_void_p),
                ("matched_rel_pos", C.c_void_p), ("insert_size", C.c_void_p), ("chr_id", C.c_void_p),
                ("close_info", C.c_void_p), ("regions", C.c_void_p), ("region_off", C.c_void_p)]


class _Results(C.Structure):
    _fields_ = [("out", C.c_void_p), ("runs", C.c_void_p), ("runs_cap", C.c_uint32), ("n_runs", C.c_uint32)]


class _SvReads(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("seq_blob", C.c_void_p), ("seq_off", C.c_void_p), ("matched_d", C.c_void_p),
                ("insert_size", C.c_void_p), ("rflags", C.c_void_p), ("run_off", C.c_void_p), ("n_close_runs", C.c_void_p),
                ("runs", C.c_void_p)]


class _SvWindow(C.Structure):
    _fields_ = [("driver_mask", C.c_uint32), ("chr", C.c_int32), ("window_end", C.c_uint32), ("num_boxes", C.c_uint32),
                ("box_size", C.c_uint32), ("region_start_defined", C.c_uint32), ("region_start", C.c_uint32),
                ("region_end_defined", C.c_uint32), ("region_end", C.c_uint32), ("min_inversion_size", C.c_int32),
                ("min_matched_bases", C.c_int32), ("reserved", C.c_int32)]


class _SvResults(C.Structure):
    _fields_ = [("records", C.c_void_p), ("records_cap", C.c_uint32), ("n_records", C.c_uint32)]


class PgError(RuntimeError):
    pass


_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PgError("libpindel_pg.so is not built (run `python -m pindel_b200.build`); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        lib.pg_create.restype = C.c_void_p
        lib.pg_create.argtypes = [C.POINTER(Settings)]
        lib.pg_create_error.restype = C.c_char_p
        lib.pg_last_error.restype = C.c_char_p
        lib.pg_last_error.argtypes = [C.c_void_p]
        lib.pg_destroy.argtypes = [C.c_void_p]
        lib.pg_kernel_launches.restype = C.c_uint64
        lib.pg_kernel_launches.argtypes = [C.c_void_p]
        lib.pg_last_run_count.restype = C.c_uint32
        lib.pg_last_run_count.argtypes = [C.c_void_p]
        lib.pg_host_alloc.restype = C.c_void_p
        lib.pg_host_alloc.argtypes = [C.c_size_t]
        lib.pg_host_free.argtypes = [C.c_void_p]
        for f in ("pg_set_stream", "pg_max_mismatch_table"):
            getattr(lib, f).argtypes = [C.c_void_p, C.c_void_p]
        lib.pg_set_reference.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        for f in ("pg_close_end", "pg_far_end", "pg_search"):
            getattr(lib, f).argtypes = [C.c_void_p, C.POINTER(_Reads), C.POINTER(_Results)]
        lib.pg_upload.argtypes = [C.c_void_p, C.POINTER(_Reads)]
        lib.pg_launch.argtypes = [C.c_void_p, C.c_int32]
        lib.pg_download.argtypes = [C.c_void_p, C.POINTER(_Results)]
        lib.pg_sync.argtypes = [C.c_void_p]
        lib.pg_last_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pg_classify.argtypes = [C.c_void_p, C.POINTER(_SvReads), C.POINTER(_SvWindow), C.POINTER(_SvResults)]
        lib.pg_classify_resident.argtypes = [C.c_void_p, C.POINTER(_SvWindow), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_SvResults)]
        lib.pg_bgzf_inflated_size.restype = C.c_int64
        lib.pg_bgzf_inflated_size.argtypes = [C.c_void_p, C.c_size_t]
        lib.pg_bgzf_inflate.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_int32]
        lib.pg_bgzf_device_data.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        lib.pg_bgzf_launches.restype = C.c_uint64
        lib.pg_bgzf_launches.argtypes = [C.c_void_p]
        lib.pg_bgzf_last_kernel_ms.restype = C.c_double
        lib.pg_bgzf_last_kernel_ms.argtypes = [C.c_void_p]
        lib.pg_bam_index.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
        lib.pg_bam_index_run.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.pg_bam_rp_discovery.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p,
                                            C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.pg_bam_sr_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p, C.c_uint64,
                                             C.POINTER(C.c_uint32), C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.pg_bam_sr_close_end.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p, C.c_uint64,
                                            C.POINTER(C.c_uint32), C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_void_p]
        lib.pg_bam_sr_close_end.restype = C.c_int
        lib.pg_bam_launches.restype = C.c_uint64
        lib.pg_bam_launches.argtypes = [C.c_void_p]
        lib.pg_bam_last_ms.restype = C.c_double
        lib.pg_bam_last_ms.argtypes = [C.c_void_p]
        lib.pg_bam_last_segments.restype = None
        lib.pg_bam_last_segments.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pg_box_sort.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pg_box_sort_kind.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pg_box_sort_launches.restype = C.c_uint64
        lib.pg_box_sort_launches.argtypes = [C.c_void_p]
        lib.pg_box_sort_last_kernel_ms.restype = C.c_double
        lib.pg_box_sort_last_kernel_ms.argtypes = [C.c_void_p]
        lib.pg_rp_modify.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        lib.pg_rp_summarize.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        lib.pg_rp_last_classes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pg_rp_last_classes.restype = None
        lib.pg_rp_modify_interchr.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pg_rp_summarize_interchr.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pg_rp_launches.restype = C.c_uint64
        lib.pg_rp_launches.argtypes = [C.c_void_p]
        lib.pg_rp_last_kernel_ms.restype = C.c_double
        lib.pg_rp_last_kernel_ms.argtypes = [C.c_void_p]
        lib.pg_classify_launches.restype = C.c_uint64
        lib.pg_classify_launches.argtypes = [C.c_void_p]
        lib.pg_classify_last_kernel_ms.restype = C.c_double
        lib.pg_classify_last_kernel_ms.argtypes = [C.c_void_p]
        lib.pg_compress_ups.restype = C.c_uint32
        lib.pg_compress_ups.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        lib.pg_device_results.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pg_device_results_previous.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pg_device_results_previous.restype = C.c_int
        lib.pg_search_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        lib.pg_search_resident.restype = C.c_int
        lib.pg_bd_load_region.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int32, C.c_uint32, C.c_uint32]
        lib.pg_bd_unload.argtypes = [C.c_void_p]
        lib.pg_bd_clusters.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        lib.pg_bd_launches.restype = C.c_uint64
        lib.pg_bd_launches.argtypes = [C.c_void_p]
        lib.pg_bd_lookups.restype = C.c_uint64
        lib.pg_bd_lookups.argtypes = [C.c_void_p]
        lib.pg_group_create.restype = C.c_void_p
        lib.pg_group_create.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        lib.pg_group_destroy.argtypes = [C.c_void_p]
        lib.pg_group_destroy.restype = None
        lib.pg_group_last_error.restype = C.c_char_p
        lib.pg_group_last_error.argtypes = [C.c_void_p]
        lib.pg_group_set_reference.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        for f in ("pg_group_close_end", "pg_group_far_end", "pg_group_search"):
            getattr(lib, f).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            getattr(lib, f).restype = C.c_int
        _lib = lib
    return _lib


def pinned_empty(n: int, dtype) -> np.ndarray:
    """numpy array backed by pinned host memory (pg_host_alloc); freed when the array is collected."""
    lib = load_library()
    dt = np.dtype(dtype)
    nbytes = max(1, n * dt.itemsize)
    p = lib.pg_host_alloc(nbytes)
    if not p:
        raise PgError("pg_host_alloc failed")
    buf = (C.c_uint8 * nbytes).from_address(p)
    arr = np.frombuffer(buf, dtype=dt, count=n)
    _PINNED[id(buf)] = (buf, p)
    import weakref
    weakref.finalize(arr, _free_pinned, id(buf))
    return arr


_PINNED = {}


def _free_pinned(key):
    ent = _PINNED.pop(key, None)
    if ent and _lib is not None:
        _lib.pg_host_free(ent[1])


class PatternGrowth:
    """One search context on one GPU (pg_ctx)."""

    def __init__(self, device: int = 0, **settings):
        self.lib = load_library()
        s = Settings()
        self.lib.pg_default_settings(C.byref(s))
        s.device = device
        for k, v in settings.items():
            if not hasattr(s, k):
                raise TypeError("unknown setting %r" % k)
            setattr(s, k, v)
        self.settings = s
        self.ctx = self.lib.pg_create(C.byref(s))
        if not self.ctx:
            raise PgError("pg_create: " + self.lib.pg_create_error().decode())
        self._keep = []
        self._n = 0

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.pg_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != PG_OK:
            raise PgError("%s failed (%d): %s" % (what, rc, self.lib.pg_last_error(self.ctx).decode()))

    def set_stream(self, cuda_stream_handle: int):
        self._check(self.lib.pg_set_stream(self.ctx, C.c_void_p(cuda_stream_handle)), "pg_set_stream")

    def max_mismatch_table(self) -> np.ndarray:
        out = np.zeros(500, dtype=np.uint32)
        self._check(self.lib.pg_max_mismatch_table(self.ctx, out.ctypes.data_as(C.c_void_p)), "pg_max_mismatch_table")
        return out

    def set_reference(self, chr_seqs):
        """chr_seqs: list of padded chromosome strings as uint8 ASCII arrays (or bytes)."""
        arrs = [np.frombuffer(c, dtype=np.uint8) if isinstance(c, (bytes, bytearray)) else
                np.ascontiguousarray(c, dtype=np.uint8) for c in chr_seqs]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        lens = np.array([a.size for a in arrs], dtype=np.uint32)
        self._check(self.lib.pg_set_reference(self.ctx, len(arrs), ptrs, lens.ctypes.data_as(C.c_void_p)), "pg_set_reference")
        self.n_chr = len(arrs)

    # ---- batches -------------------------------------------------------------------------------
    def _reads(self, seq_blob, seq_off, matched_d=None, matched_rel_pos=None, insert_size=None, chr_id=None,
               close_info=None, regions=None, region_off=None):
        def arr(a, dt):
            return None if a is None else np.ascontiguousarray(a, dtype=dt)
        keep = dict(seq_blob=arr(seq_blob, np.uint8), seq_off=arr(seq_off, np.uint32), matched_d=arr(matched_d, np.uint8),
                    matched_rel_pos=arr(matched_rel_pos, np.uint32), insert_size=arr(insert_size, np.int16),
                    chr_id=arr(chr_id, np.int32), close_info=arr(close_info, R.CLOSEINFO_DT),
                    regions=arr(regions, R.REGION_DT), region_off=arr(region_off, np.uint32))
        r = _Reads()
        r.n = len(keep["seq_off"]) - 1
        for k, v in keep.items():
            setattr(r, k, None if v is None else v.ctypes.data)
        return r, keep

    def upload(self, **batch):
        r, keep = self._reads(**batch)
        self._keep = keep
        self._n = r.n
        self._check(self.lib.pg_upload(self.ctx, C.byref(r)), "pg_upload")

    def launch(self, mode: int = 3):
        self._check(self.lib.pg_launch(self.ctx, mode), "pg_launch")

    def sync(self):
        self._check(self.lib.pg_sync(self.ctx), "pg_sync")

    def download(self, out=None, runs=None):
        n = self._n
        out = np.zeros(n, dtype=R.OUT_DT) if out is None else out
        runs = np.zeros(max(1024, 24 * n), dtype=R.RUN_DT) if runs is None else runs
        res = _Results(out.ctypes.data, runs.ctypes.data, runs.size, 0)
        rc = self.lib.pg_download(self.ctx, C.byref(res))
        if rc == PG_ERR_RUNS_CAP:
            runs = np.zeros(res.n_runs, dtype=R.RUN_DT)
            res = _Results(out.ctypes.data, runs.ctypes.data, runs.size, 0)
            rc = self.lib.pg_download(self.ctx, C.byref(res))
        self._check(rc, "pg_download")
        return out, runs[:res.n_runs]

    def _call(self, fn, what, out=None, runs=None, **batch):
        r, keep = self._reads(**batch)
        n = r.n
        self._n = n
        out = np.zeros(n, dtype=R.OUT_DT) if out is None else out
        runs = np.empty(max(1024, 8 * n), dtype=R.RUN_DT) if runs is None else runs
        res = _Results(out.ctypes.data, runs.ctypes.data, runs.size, 0)
        rc = fn(self.ctx, C.byref(r), C.byref(res))
        if rc == PG_ERR_RUNS_CAP:        # the required size came back in n_runs: repeat the call with room for it
            runs = np.empty(res.n_runs, dtype=R.RUN_DT)
            res = _Results(out.ctypes.data, runs.ctypes.data, runs.size, 0)
            rc = fn(self.ctx, C.byref(r), C.byref(res))
        self._check(rc, what)
        return out, runs[:res.n_runs]

    def get_close_end(self, **batch):
        """Batched GetCloseEnd (src/pindel.cpp:2531-2605): -> (out[n], runs); UP_Close before CleanUniquePoints."""
        return self._call(self.lib.pg_close_end, "pg_close_end", **batch)

    def search_far_ends(self, **batch):
        """Batched SearchFarEnds (src/pindel.cpp:1115-1138); needs close_info, the post-flip sequences, optional clusters."""
        return self._call(self.lib.pg_far_end, "pg_far_end", **batch)

    def search(self, **batch):
        """GetCloseEnd + SearchFarEnd per read, fused on the device."""
        return self._call(self.lib.pg_search, "pg_search", **batch)

    def last_stats(self):
        """(reference positions scanned, UniquePoints produced) by the last launch; waits for it."""
        wb, up = C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.pg_last_stats(self.ctx, C.byref(wb), C.byref(up)), "pg_last_stats")
        return int(wb.value), int(up.value)

    def device_results(self):
        """(out device pointer, runs device pointer, n_runs) of the last launch; waits for it."""
        o, r, k = C.c_void_p(0), C.c_void_p(0), C.c_uint32(0)
        self._check(self.lib.pg_device_results(self.ctx, C.byref(o), C.byref(r), C.byref(k)), "pg_device_results")
        return int(o.value or 0), int(r.value or 0), int(k.value)

    def device_results_previous(self):
        """(out device pointer, runs device pointer, n_runs, n_reads) of the launch before the last one; waits for that
        launch only, so the latest launch keeps running while these records are handed on."""
        o, r, k, n = C.c_void_p(0), C.c_void_p(0), C.c_uint32(0), C.c_int32(0)
        self._check(self.lib.pg_device_results_previous(self.ctx, C.byref(o), C.byref(r), C.byref(k), C.byref(n)),
                    "pg_device_results_previous")
        return int(o.value or 0), int(r.value or 0), int(k.value), int(n.value)

    def search_resident(self, mode: int = 3, **batch):
        """pg_upload + pg_launch with the upload chunked behind the kernel; batch and results stay on the device.  The
        arrays are kept alive by this object until the next batch call."""
        r, keep = self._reads(**batch)
        self._keep = keep
        self._n = r.n
        self._check(self.lib.pg_search_resident(self.ctx, C.byref(r), mode), "pg_search_resident")

    def classify(self, batch: dict, win: dict, records=None) -> np.ndarray:
        """The seven SV drivers of SearchSVs (src/pindel.cpp:1140-1164) on one window: batch = the arrays of pg_sv_reads
        (pindel_b200.sv.batch_from_lists), win = pindel_b200.sv.window(...).  -> records (sv.RECORD_DT) sorted by
        (driver, read, seq); replay them with pindel_b200.sv.replay."""
        from . import sv
        n = len(batch["matched_d"])
        arrs = {k: np.ascontiguousarray(batch[k], dtype=dt) for k, dt in
                (("seq_blob", np.uint8), ("seq_off", np.uint32), ("matched_d", np.uint8), ("insert_size", np.int16),
                 ("rflags", np.uint8), ("run_off", np.uint32), ("n_close_runs", np.uint16), ("runs", R.RUN_DT))}
        r = _SvReads(n, *[arrs[k].ctypes.data for k in ("seq_blob", "seq_off", "matched_d", "insert_size", "rflags", "run_off",
                                                        "n_close_runs", "runs")])
        reg = win.get("region")
        w = _SvWindow(win["driver_mask"], win.get("chr", 0), win["window_end"], win["num_boxes"], win["box_size"],
                      0 if reg is None else 1, 0 if reg is None else reg[0], 0 if reg is None else 1, 0 if reg is None else reg[1],
                      win["min_inversion_size"], win["min_matched_bases"], 0)
        rec = np.empty(max(1024, 2 * n), dtype=sv.RECORD_DT) if records is None else records
        res = _SvResults(rec.ctypes.data, rec.size, 0)
        rc = self.lib.pg_classify(self.ctx, C.byref(r), C.byref(w), C.byref(res))
        if rc == PG_ERR_RUNS_CAP:
            rec = np.empty(res.n_records, dtype=sv.RECORD_DT)
            res = _SvResults(rec.ctypes.data, rec.size, 0)
            rc = self.lib.pg_classify(self.ctx, C.byref(r), C.byref(w), C.byref(res))
        self._check(rc, "pg_classify")
        return rec[:res.n_records]

    def classify_resident(self, win: dict, rflags=None, close_run_off=None, close_runs=None, records=None) -> np.ndarray:
        """pg_classify on the batch that is still on the device after search() / search_far_ends() / launch(2|3): nothing
        but the window parameters (and, after a far-end-only launch, the reads' close runs) is uploaded."""
        from . import sv
        n = self._n
        reg = win.get("region")
        w = _SvWindow(win["driver_mask"], win.get("chr", 0), win["window_end"], win["num_boxes"], win["box_size"],
                      0 if reg is None else 1, 0 if reg is None else reg[0], 0 if reg is None else 1, 0 if reg is None else reg[1],
                      win["min_inversion_size"], win["min_matched_bases"], 0)
        fl = None if rflags is None else np.ascontiguousarray(rflags, dtype=np.uint8)
        co = None if close_run_off is None else np.ascontiguousarray(close_run_off, dtype=np.uint32)
        cr = None if close_runs is None else np.ascontiguousarray(close_runs, dtype=R.RUN_DT)
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        rec = np.empty(max(1024, 2 * n), dtype=sv.RECORD_DT) if records is None else records
        res = _SvResults(rec.ctypes.data, rec.size, 0)
        rc = self.lib.pg_classify_resident(self.ctx, C.byref(w), ptr(fl), ptr(co), ptr(cr), C.byref(res))
        if rc == PG_ERR_RUNS_CAP:
            rec = np.empty(res.n_records, dtype=sv.RECORD_DT)
            res = _SvResults(rec.ctypes.data, rec.size, 0)
            rc = self.lib.pg_classify_resident(self.ctx, C.byref(w), ptr(fl), ptr(co), ptr(cr), C.byref(res))
        self._check(rc, "pg_classify_resident")
        return rec[:res.n_records]

    def bgzf_inflate(self, data, to_host: bool = True, check_crc: bool = True, out=None):
        """All BGZF blocks of `data` (bytes or uint8 array) inflated on the device.  -> uint8 array (to_host) or the
        inflated size (the stream then stays on the device: bgzf_device_data())."""
        buf = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else np.ascontiguousarray(data, dtype=np.uint8)
        n = C.c_size_t(0)
        if not to_host:
            self._check(self.lib.pg_bgzf_inflate(self.ctx, buf.ctypes.data, buf.size, None, 0, C.byref(n), int(check_crc)), "pg_bgzf_inflate")
            return int(n.value)
        if out is None:
            total = int(self.lib.pg_bgzf_inflated_size(buf.ctypes.data, buf.size))
            if total < 0:
                raise PgError("pg_bgzf_inflated_size: not a BGZF file")
            out = np.empty(total, dtype=np.uint8)
        self._check(self.lib.pg_bgzf_inflate(self.ctx, buf.ctypes.data, buf.size, out.ctypes.data, out.size, C.byref(n), int(check_crc)),
                    "pg_bgzf_inflate")
        return out[:n.value]

    def bgzf_device_data(self):
        p, n = C.c_void_p(0), C.c_size_t(0)
        self._check(self.lib.pg_bgzf_device_data(self.ctx, C.byref(p), C.byref(n)), "pg_bgzf_device_data")
        return int(p.value or 0), int(n.value)

    @property
    def bgzf_last_kernel_ms(self) -> float:
        return float(self.lib.pg_bgzf_last_kernel_ms(self.ctx))

    @property
    def bgzf_launches(self) -> int:
        return int(self.lib.pg_bgzf_launches(self.ctx))

    def bam_index(self, recs=None):
        """Record table of the inflated stream the last bgzf_inflate left on the device -> (records (pindel_b200.bam.REC_DT),
        first record offset, n_ref)."""
        from . import bam
        n, first, nref = C.c_uint64(0), C.c_uint64(0), C.c_int32(0)
        if recs is None:
            rc = self.lib.pg_bam_index(self.ctx, None, 0, C.byref(n), C.byref(first), C.byref(nref))
            if rc not in (0, PG_ERR_RUNS_CAP):
                self._check(rc, "pg_bam_index")
            recs = np.zeros(int(n.value), dtype=bam.REC_DT)
            if n.value == 0:
                return recs, int(first.value), int(nref.value)
        self._check(self.lib.pg_bam_index(self.ctx, recs.ctypes.data, recs.size, C.byref(n), C.byref(first), C.byref(nref)), "pg_bam_index")
        return recs[:n.value], int(first.value), int(nref.value)

    def bam_index_run(self, first_record: int, n_ref: int):
        """Record table of a headerless run of BGZF blocks left on the device by bgzf_inflate (a window's chunk): the first
        record starts at first_record, the last one may be cut off.  -> (records, tail offset)"""
        from . import bam
        n, tail = C.c_uint64(0), C.c_uint64(0)
        rc = self.lib.pg_bam_index_run(self.ctx, C.c_uint64(first_record), n_ref, None, C.c_uint64(0), C.byref(n), C.byref(tail))
        if rc not in (0, PG_ERR_RUNS_CAP):
            self._check(rc, "pg_bam_index_run")
        recs = np.zeros(int(n.value), dtype=bam.REC_DT)
        if recs.size:
            self._check(self.lib.pg_bam_index_run(self.ctx, C.c_uint64(first_record), n_ref, recs.ctypes.data, C.c_uint64(recs.size), C.byref(n), C.byref(tail)),
                        "pg_bam_index_run")
        return recs, int(tail.value)

    def bam_rp_discovery(self, tid, beg, end, insert_size, min_anchor_quality=0):
        """Discordant-pair discovery on the indexed stream (after bgzf_inflate + bam_index).
        -> (rows (rp.RP_DT), src, inter_rows, inter_src, records delivered by the region iterator)"""
        from . import rp
        q = np.array([(tid, beg, end, insert_size, min_anchor_quality)],
                     dtype=[("tid", "<i4"), ("beg", "<i4"), ("end", "<i4"), ("insert_size", "<i4"), ("min_anchor_quality", "<u4")])
        n, ni, ns = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        rc = self.lib.pg_bam_rp_discovery(self.ctx, q.ctypes.data, None, None, 0, C.byref(n), None, None, 0, C.byref(ni), C.byref(ns))
        if rc not in (0, PG_ERR_RUNS_CAP):
            self._check(rc, "pg_bam_rp_discovery")
        rows, src = np.zeros(int(n.value), dtype=rp.RP_DT), np.zeros(int(n.value), dtype=np.uint32)
        irows, isrc = np.zeros(int(ni.value), dtype=rp.RP_DT), np.zeros(int(ni.value), dtype=np.uint32)
        if rc == PG_ERR_RUNS_CAP:
            vp = lambda a: a.ctypes.data_as(C.c_void_p) if a.size else None
            self._check(self.lib.pg_bam_rp_discovery(self.ctx, q.ctypes.data, vp(rows), vp(src), rows.size, C.byref(n), vp(irows), vp(isrc), irows.size,
                                                     C.byref(ni), C.byref(ns)), "pg_bam_rp_discovery")
        return rows, src, irows, isrc, int(ns.value)

    def bam_sr_candidates(self, tid, beg, end, insert_size, chr_len, min_anchor_quality=0, nm_limit=2, max_mismatch_rate=0.02):
        """The reader's split-read filter on the indexed stream (after bgzf_inflate + bam_index): what ReadInBamReads_SR hands
        to its ReadBuffer.  -> (cands (bam.SR_CAND_DT), seqs [n, stride] uint8, refs (bam.SR_REF_DT), records delivered)"""
        from . import bam
        q = np.array([(tid, beg, end, insert_size, min_anchor_quality, chr_len, nm_limit, 0, max_mismatch_rate)], dtype=bam.SR_QUERY_DT)
        n, m, ns, stride = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
        rc = self.lib.pg_bam_sr_candidates(self.ctx, q.ctypes.data, None, C.c_uint64(0), C.byref(n), None, C.c_uint64(0), C.byref(stride), None, C.c_uint64(0),
                                           C.byref(m), C.byref(ns))
        if rc not in (0, PG_ERR_RUNS_CAP):
            self._check(rc, "pg_bam_sr_candidates")
        cands = np.zeros(int(n.value), dtype=bam.SR_CAND_DT)
        seqs = np.zeros((int(n.value), max(int(stride.value), 1)), dtype=np.uint8)
        refs = np.zeros(int(m.value), dtype=bam.SR_REF_DT)
        if rc == PG_ERR_RUNS_CAP:
            vp = lambda a: a.ctypes.data_as(C.c_void_p) if a.size else None
            self._check(self.lib.pg_bam_sr_candidates(self.ctx, q.ctypes.data, vp(cands), C.c_uint64(cands.size), C.byref(n), vp(seqs), C.c_uint64(seqs.size),
                                                      C.byref(stride), vp(refs), C.c_uint64(refs.size), C.byref(m), C.byref(ns)), "pg_bam_sr_candidates")
        return cands, seqs, refs, int(ns.value)

    def bam_sr_close_end(self, tid, beg, end, insert_size, chr_len, chr_id=0, min_anchor_quality=0, nm_limit=2, max_mismatch_rate=0.02):
        """bam_sr_candidates with the candidates searched where they lie (GetCloseEnd on the device, no sequence upload).
        -> (cands, seqs [n, stride], refs, records delivered, out (runs.OUT_DT), runs (runs.RUN_DT))"""
        from . import bam
        q = np.array([(tid, beg, end, insert_size, min_anchor_quality, chr_len, nm_limit, 0, max_mismatch_rate)], dtype=bam.SR_QUERY_DT)
        n, m, ns, stride = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
        res = _Results(None, None, 0, 0)
        rc = self.lib.pg_bam_sr_close_end(self.ctx, q.ctypes.data, chr_id, None, C.c_uint64(0), C.byref(n), None, C.c_uint64(0), C.byref(stride), None,
                                          C.c_uint64(0), C.byref(m), C.byref(ns), C.byref(res))
        if rc not in (0, PG_ERR_RUNS_CAP):
            self._check(rc, "pg_bam_sr_close_end")
        cands = np.zeros(int(n.value), dtype=bam.SR_CAND_DT)
        seqs = np.zeros((int(n.value), max(int(stride.value), 1)), dtype=np.uint8)
        refs = np.zeros(int(m.value), dtype=bam.SR_REF_DT)
        out = np.zeros(int(n.value), dtype=R.OUT_DT)
        runs = np.zeros(max(1, int(res.n_runs)), dtype=R.RUN_DT)
        if rc == PG_ERR_RUNS_CAP:
            vp = lambda a: a.ctypes.data_as(C.c_void_p) if a.size else None
            res = _Results(out.ctypes.data, runs.ctypes.data, runs.size, 0)
            self._check(self.lib.pg_bam_sr_close_end(self.ctx, q.ctypes.data, chr_id, vp(cands), C.c_uint64(cands.size), C.byref(n), vp(seqs),
                                                     C.c_uint64(seqs.size), C.byref(stride), vp(refs), C.c_uint64(refs.size), C.byref(m), C.byref(ns),
                                                     C.byref(res)), "pg_bam_sr_close_end")
        self._n = int(n.value)
        return cands, seqs, refs, int(ns.value), out, runs[:res.n_runs]

    # ---- BreakDancer mask (row N2, last part) ----------------------------------------------------
    def bd_load_region(self, events, chr_name_rank, chr_id, start, end):
        """BDData::loadRegion for SearchWindow(chr_id, start, end): events [n, 6] int (chr_a, pos_a, pos2_a, chr_b, pos_b, pos2_b)
        in the reference's order; chr_name_rank[c] = rank of chromosome c's name in string order."""
        ev = np.ascontiguousarray(np.array(events, dtype=np.int64).reshape(-1, 6).astype(np.uint32).view(np.int32))
        rk = np.ascontiguousarray(chr_name_rank, dtype=np.int32)
        self._check(self.lib.pg_bd_load_region(self.ctx, ev.ctypes.data_as(C.c_void_p), ev.shape[0], rk.ctypes.data_as(C.c_void_p), chr_id,
                                               C.c_uint32(start), C.c_uint32(end)), "pg_bd_load_region")

    def bd_unload(self):
        self._check(self.lib.pg_bd_unload(self.ctx), "pg_bd_unload")

    def bd_clusters(self, positions):
        """getCorrespondingSearchWindowCluster for every close-end position -> (region_off [n+1], regions (runs.REGION_DT))"""
        q = np.ascontiguousarray(positions, dtype=np.uint32)
        off = np.zeros(q.size + 1, dtype=np.uint32)
        n = C.c_uint64(0)
        rc = self.lib.pg_bd_clusters(self.ctx, q.size, q.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), None, C.c_uint64(0), C.byref(n))
        if rc not in (0, PG_ERR_RUNS_CAP):
            self._check(rc, "pg_bd_clusters")
        reg = np.zeros(int(n.value), dtype=R.REGION_DT)
        if rc == PG_ERR_RUNS_CAP:
            self._check(self.lib.pg_bd_clusters(self.ctx, q.size, q.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                                                reg.ctypes.data_as(C.c_void_p), C.c_uint64(reg.size), C.byref(n)), "pg_bd_clusters")
        return off, reg

    @property
    def bd_lookups(self) -> int:
        return int(self.lib.pg_bd_lookups(self.ctx))

    def bam_last_segments(self):
        a, b = C.c_uint32(0), C.c_uint32(0)
        self.lib.pg_bam_last_segments(self.ctx, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    @property
    def bam_last_ms(self) -> float:
        return float(self.lib.pg_bam_last_ms(self.ctx))

    @property
    def bam_launches(self) -> int:
        return int(self.lib.pg_bam_launches(self.ctx))

    def box_sort(self, member_keys: np.ndarray, box_off: np.ndarray):
        """bubblesortReads + markDuplicates for all boxes at once.  member_keys: 32-byte keys per member in box order
        (include/pindel_pg.h pg_sort_key).  -> (order within each box, nonunique per sorted position)"""
        keys = np.ascontiguousarray(member_keys)
        assert keys.dtype.itemsize == 32
        box_off = np.ascontiguousarray(box_off, dtype=np.uint32)
        order = np.zeros(keys.size, dtype=np.uint32)
        non = np.zeros(keys.size, dtype=np.uint8)
        vp = lambda a: a.ctypes.data_as(C.c_void_p) if a.size else None
        self._check(self.lib.pg_box_sort(self.ctx, box_off.size - 1, vp(box_off), vp(keys), vp(order), vp(non)), "pg_box_sort")
        return order, non

    def box_sort_kind(self, kind: int, member_keys: np.ndarray, box_off: np.ndarray):
        """The ordering, duplicate marks and event cut of one reporter (include/pindel_pg.h PG_BOX_*) for all boxes at once.
        -> (order within each box, nonunique per sorted position, event flags per sorted position)"""
        keys = np.ascontiguousarray(member_keys)
        assert keys.dtype.itemsize == 32
        box_off = np.ascontiguousarray(box_off, dtype=np.uint32)
        order = np.zeros(keys.size, dtype=np.uint32)
        non = np.zeros(keys.size, dtype=np.uint8)
        ev = np.zeros(keys.size, dtype=np.uint8)
        vp = lambda a: a.ctypes.data_as(C.c_void_p) if a.size else None
        self._check(self.lib.pg_box_sort_kind(self.ctx, kind, box_off.size - 1, vp(box_off), vp(keys), vp(order), vp(non), vp(ev)), "pg_box_sort_kind")
        return order, non, ev

    @property
    def box_sort_last_kernel_ms(self) -> float:
        return float(self.lib.pg_box_sort_last_kernel_ms(self.ctx))

    @property
    def box_sort_launches(self) -> int:
        return int(self.lib.pg_box_sort_launches(self.ctx))

    def rp_modify(self, rows: np.ndarray) -> np.ndarray:
        """ModifyRP after its sort (src/bddata.cpp:396-430) on the device: rows = pindel_b200.rp.RP_DT in the order the
        caller's Compare2RP sort leaves -> pindel_b200.rp.CLUSTER_DT."""
        from . import rp
        rows = np.ascontiguousarray(rows, dtype=rp.RP_DT)
        out = np.zeros(rows.size, dtype=rp.CLUSTER_DT)
        self._check(self.lib.pg_rp_modify(self.ctx, int(rows.size), rows.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)),
                    "pg_rp_modify")
        return out

    def rp_summarize(self, rows: np.ndarray, state: np.ndarray) -> np.ndarray:
        """Summarize (src/bddata.cpp:470-575) on the device; returns the updated copy of `state`."""
        from . import rp
        rows = np.ascontiguousarray(rows, dtype=rp.RP_DT)
        out = np.ascontiguousarray(state, dtype=rp.CLUSTER_DT).copy()
        self._check(self.lib.pg_rp_summarize(self.ctx, int(rows.size), rows.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)),
                    "pg_rp_summarize")
        return out

    def rp_cluster_interchr(self, rows: np.ndarray, chr_a, chr_b):
        """ModifyRP_InterChr (after its sort) + Summarize_InterChr (src/bddata.cpp:439-468, 577-624) on the device.
        -> (state after modify, state after summarize)"""
        from . import rp
        rows = np.ascontiguousarray(rows, dtype=rp.RP_DT)
        ca, cb = np.ascontiguousarray(chr_a, dtype=np.int32), np.ascontiguousarray(chr_b, dtype=np.int32)
        out = np.zeros(rows.size, dtype=rp.CLUSTER_DT)
        vp = lambda a: a.ctypes.data_as(C.c_void_p) if a.size else None
        self._check(self.lib.pg_rp_modify_interchr(self.ctx, int(rows.size), vp(rows), vp(ca), vp(cb), vp(out)), "pg_rp_modify_interchr")
        mod = out.copy()
        self._check(self.lib.pg_rp_summarize_interchr(self.ctx, int(rows.size), vp(rows), vp(ca), vp(cb), vp(out)), "pg_rp_summarize_interchr")
        return mod, out

    def rp_last_classes(self):
        """(classes run by one warp, classes run by one CTA, rows in either) of the last rp_modify."""
        a, b, c = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        self.lib.pg_rp_last_classes(self.ctx, C.byref(a), C.byref(b), C.byref(c))
        return int(a.value), int(b.value), int(c.value)

    @property
    def rp_last_kernel_ms(self) -> float:
        return float(self.lib.pg_rp_last_kernel_ms(self.ctx))

    @property
    def rp_launches(self) -> int:
        return int(self.lib.pg_rp_launches(self.ctx))

    @property
    def classify_last_kernel_ms(self) -> float:
        return float(self.lib.pg_classify_last_kernel_ms(self.ctx))

    @property
    def classify_launches(self) -> int:
        return int(self.lib.pg_classify_launches(self.ctx))

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.pg_kernel_launches(self.ctx))

    @property
    def last_run_count(self) -> int:
        return int(self.lib.pg_last_run_count(self.ctx))


class PatternGrowthGroup:
    """Several GPUs behind one handle (pg_group): the batch calls of PatternGrowth with the reads dealt to the devices in
    contiguous ranges inside the library; results come back in the caller's order."""

    def __init__(self, devices, **settings):
        self.lib = load_library()
        s = Settings()
        self.lib.pg_default_settings(C.byref(s))
        for k, v in settings.items():
            if not hasattr(s, k):
                raise TypeError("unknown setting %r" % k)
            setattr(s, k, v)
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        self.grp = self.lib.pg_group_create(C.byref(s), int(dev.size), dev.ctypes.data_as(C.c_void_p))
        if not self.grp:
            raise PgError("pg_group_create: " + self.lib.pg_create_error().decode())
        self._reads = PatternGrowth._reads.__get__(self)

    def close(self):
        if getattr(self, "grp", None):
            self.lib.pg_group_destroy(self.grp)
            self.grp = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != PG_OK:
            raise PgError("%s failed (%d): %s" % (what, rc, self.lib.pg_group_last_error(self.grp).decode()))

    def set_reference(self, chr_seqs):
        arrs = [np.frombuffer(c, dtype=np.uint8) if isinstance(c, (bytes, bytearray)) else
                np.ascontiguousarray(c, dtype=np.uint8) for c in chr_seqs]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        lens = np.array([a.size for a in arrs], dtype=np.uint32)
        self._check(self.lib.pg_group_set_reference(self.grp, len(arrs), ptrs, lens.ctypes.data_as(C.c_void_p)), "pg_group_set_reference")

    def _call(self, fn, what, out=None, runs=None, **batch):
        r, keep = self._reads(**batch)
        n = r.n
        out = np.zeros(n, dtype=R.OUT_DT) if out is None else out
        runs = np.empty(max(1024, 8 * n), dtype=R.RUN_DT) if runs is None else runs
        res = _Results(out.ctypes.data, runs.ctypes.data, runs.size, 0)
        rc = fn(self.grp, C.byref(r), C.byref(res))
        if rc == PG_ERR_RUNS_CAP:
            runs = np.empty(res.n_runs, dtype=R.RUN_DT)
            res = _Results(out.ctypes.data, runs.ctypes.data, runs.size, 0)
            rc = fn(self.grp, C.byref(r), C.byref(res))
        self._check(rc, what)
        return out, runs[:res.n_runs]

    def get_close_end(self, **batch):
        return self._call(self.lib.pg_group_close_end, "pg_group_close_end", **batch)

    def search_far_ends(self, **batch):
        return self._call(self.lib.pg_group_far_end, "pg_group_far_end", **batch)

    def search(self, **batch):
        return self._call(self.lib.pg_group_search, "pg_group_search", **batch)
