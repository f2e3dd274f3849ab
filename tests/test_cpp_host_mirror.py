"""include/pindel_pg.hpp -- the header-only C++ mirror of the reference's seam (GetCloseEnd / SearchFarEnds /
CleanUniquePoints / ReverseComplement on SPLIT_READ vectors): it must compile with plain g++ against the C ABI, refuse to
run without a device, and on the GPU return what the oracle returns."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import refhook
from pindel_b200 import synth
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")
LIBDIR = os.path.join(ROOT, "pindel_b200")


def build_exe():
    deps = [SRC, os.path.join(ROOT, "include", "pindel_pg.hpp"), os.path.join(ROOT, "include", "pindel_pg.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++11", "-Wall", "-Wextra", "-o", EXE, SRC, "-L" + LIBDIR, "-lpindel_pg",
                               "-Wl,-rpath," + LIBDIR])
    return EXE


def write_input(path, chroms, batch, chr_id):
    with open(path, "wb") as f:
        f.write(struct.pack("<i", len(chroms)))
        for c in chroms:
            f.write(struct.pack("<I", c.size)); f.write(c.tobytes())
        f.write(struct.pack("<i", batch.n))
        for i in range(batch.n):
            s = batch.seq[i].tobytes()
            f.write(struct.pack("<i", len(s))); f.write(s)
            f.write(struct.pack("<bIhi", int(batch.matched_d[i]), int(batch.matched_rel_pos[i]), int(batch.insert_size[i]), int(chr_id[i])))


def test_header_compiles_with_plain_gxx_and_fails_loudly_without_a_device(tmp_path):
    exe = build_exe()
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("device present: covered by the gpu test")
    ref = synth.make_reference(300_000, 71)
    batch = synth.make_reads(ref, 5, 100, seed=72)
    write_input(tmp_path / "in.bin", [synth.pad_reference(ref)], batch, np.zeros(5, np.int32))
    p = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert p.returncode == 1 and "no CUDA device" in p.stderr


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle(tmp_path):
    exe = build_exe()
    ref = synth.make_reference(500_000, 73, n_runs=[(200_000, 800)])
    padded = synth.pad_reference(ref)
    batch = synth.make_reads(ref, 3000, 100, seed=74)
    write_input(tmp_path / "in.bin", [padded], batch, np.zeros(batch.n, np.int32))
    p = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    got = refhook.parse_stream(np.fromfile(tmp_path / "out.bin", dtype=np.int32), batch.n)
    exp = H.oracle_results([padded], batch, threads=4)
    H.compare(exp, got, "C++ host mirror")
