"""The C-ABI shared library: it loads without a GPU, exports every symbol include/pindel_pg.h declares, refuses to
create a context when no device exists (no CPU fallback), and its pure-host helpers agree with the Python mirror."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from pindel_b200 import api, runs as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "pindel_pg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = api.load_library()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libpindel_pg.so does not export " + s


def test_struct_layouts_match_header():
    hdr = open(os.path.join(ROOT, "include", "pindel_pg.h")).read()
    assert "uint32_t abs_loc0;" in hdr and R.RUN_DT.itemsize == 12
    assert R.OUT_DT.itemsize == 12 and R.REGION_DT.itemsize == 12 and R.CLOSEINFO_DT.itemsize == 8
    assert C.sizeof(api.Settings) == 48


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a machine without a CUDA device")
def test_no_cpu_fallback_without_a_device():
    with pytest.raises(api.PgError) as e:
        api.PatternGrowth()
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_host_helpers_expand_and_clean():
    lib = api.load_library()
    lib.pg_expand_runs.restype = C.c_uint32
    lib.pg_clean_unique_points.restype = C.c_uint32
    runs = np.zeros(3, dtype=R.RUN_DT)
    runs[0] = (1000, 8, 5, 0, 2, 0)      # FORWARD/ANTISENSE, len 8..12
    runs[1] = (2000, 13, 2, 1, 2, 0)     # another seed
    runs[2] = (1008, 16, 4, 1, 2, 0)     # same terminal as run 0 (abs_loc - len == 992)
    UP = np.dtype([("abs_loc", "<u4"), ("len", "<i2"), ("mm", "<i2"), ("dir", "u1"), ("strand", "u1"), ("chr", "<i2")])
    n = lib.pg_expand_runs(runs.ctypes.data_as(C.c_void_p), 3, None)
    assert n == 11
    ups = np.zeros(n, dtype=UP)
    lib.pg_expand_runs(runs.ctypes.data_as(C.c_void_p), 3, ups.ctypes.data_as(C.c_void_p))
    exp = R.expand_runs(runs)
    assert np.array_equal(ups["len"], exp[:, 0]) and np.array_equal(ups["abs_loc"], exp[:, 1])
    assert np.array_equal(ups["dir"], exp[:, 2]) and np.array_equal(ups["strand"], exp[:, 3]) and np.array_equal(ups["mm"], exp[:, 4])
    k = lib.pg_clean_unique_points(ups.ctypes.data_as(C.c_void_p), n)
    cl = R.clean_unique_points(exp)
    assert k == cl.shape[0] == 9
    assert np.array_equal(ups["len"][:k], cl[:, 0]) and np.array_equal(ups["abs_loc"][:k], cl[:, 1])


def test_host_helper_compress_is_the_inverse_of_expand():
    lib = api.load_library()
    from pindel_b200 import sv
    runs = np.zeros(4, dtype=R.RUN_DT)
    runs[0] = (1000, 8, 5, 0, 2, 0)
    runs[1] = (2000, 13, 2, 1, 1, 0)     # BACKWARD
    runs[2] = (1998, 15, 1, 1, 1, 0)     # continues run 1 in position and length: merges with it
    runs[3] = (70000, 16, 4, 2, 3, 1)
    UP = np.dtype([("abs_loc", "<u4"), ("len", "<i2"), ("mm", "<i2"), ("dir", "u1"), ("strand", "u1"), ("chr", "<i2")])
    n = lib.pg_expand_runs(runs.ctypes.data_as(C.c_void_p), 4, None)
    ups = np.zeros(n, dtype=UP)
    lib.pg_expand_runs(runs.ctypes.data_as(C.c_void_p), 4, ups.ctypes.data_as(C.c_void_p))
    k = lib.pg_compress_ups(ups.ctypes.data_as(C.c_void_p), n, None)
    assert k == 3
    back = np.zeros(k, dtype=R.RUN_DT)
    lib.pg_compress_ups(ups.ctypes.data_as(C.c_void_p), n, back.ctypes.data_as(C.c_void_p))
    assert np.array_equal(R.expand_runs(back), R.expand_runs(runs))
    assert np.array_equal(sv.compress_ups(R.expand_runs(runs)), back)      # the numpy mirror agrees with the C helper


def test_backward_runs_walk_down():
    runs = np.zeros(1, dtype=R.RUN_DT)
    runs[0] = (5000, 10, 3, 0, 1, 0)     # BACKWARD/SENSE
    e = R.expand_runs(runs)
    assert e[:, 1].tolist() == [5000, 4999, 4998] and e[:, 0].tolist() == [10, 11, 12] and e[:, 2].tolist() == [1, 1, 1]


def test_final_sequence_replays_flips():
    assert R.final_sequence(b"ACGTN\r\n", 0) == b"ACGTN"
    assert R.final_sequence(b"AACG", 1) == b"CGTT"
    assert R.final_sequence(b"MACG", 1) == b"CGT"              # leading IUPAC -> trailing NUL -> stripped
    assert R.final_sequence(b"ACGR", 2) == b"ACG"
    assert R.final_sequence(b"AMCG", 1) == b"CG\x00T"


def test_header_is_plain_c(tmp_path):
    """the boundary is a C ABI: include/pindel_pg.h must compile as C99 (-pedantic), which is what a cgo / JNI / ctypes-style
    binding on the reference's side includes; and as C++ for the reference's own translation units"""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "pindel_pg.h"\nint main(void) { pg_settings s; (void)s; return PG_OK; }\n')
    for cc, std in (("/usr/bin/gcc", "-std=c99"), ("/usr/bin/g++", "-std=c++11")):
        lang = ["-x", "c++"] if cc.endswith("g++") else []
        subprocess.check_call([cc, std, "-pedantic", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), "-c"] + lang +
                              [str(src), "-o", str(tmp_path / "t.o")])


def test_emulated_abi_exports_what_the_bridge_binds():
    """tests/emul/abi_emul.cpp stands in for the library under the drop-in bridge: every pg_* symbol the bridge object imports
    must be declared in the header (the bridge uses nothing outside the published ABI)"""
    import subprocess
    obj = os.path.join(ROOT, "oracle", "_ref", "obj", "ref_gpu_bridge.o")
    if not os.path.exists(obj):
        pytest.skip("oracle/_ref not built")
    out = subprocess.run(["nm", "-u", obj], capture_output=True, text=True, check=True).stdout
    used = sorted(set(re.findall(r"\bU (pg_[a-z0-9_]+)\b", out)))
    assert len(used) >= 15, used
    missing = [s for s in used if s not in declared_symbols()]
    assert not missing, "the bridge binds symbols the header does not declare: %s" % missing
