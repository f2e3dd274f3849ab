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
EMUL_DIR = os.path.join(ROOT, "oracle", "_ref")


def build_exe(src=SRC, exe=EXE):
    deps = [src, os.path.join(ROOT, "include", "pindel_pg.hpp"), os.path.join(ROOT, "include", "pindel_pg.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++11", "-Wall", "-Wextra", "-o", exe, src, "-L" + LIBDIR, "-lpindel_pg",
                               "-Wl,-rpath," + LIBDIR])
    return exe


STAGES_SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_stages_test.cpp")
STAGES_EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_stages_test")


def write_stages_input(path, rp_rows, reads, box_off, idx, bam_bytes):
    with open(path, "wb") as f:
        f.write(struct.pack("<i", rp_rows.size))
        for r in rp_rows:
            f.write(struct.pack("<IIIhBB", int(r["pos_a"]), int(r["pos_b"]), int(r["insert_size"]), int(r["read_length"]), int(r["da"]), int(r["db"])))
        f.write(struct.pack("<i", reads.size))
        for r in reads:
            f.write(struct.pack("<IIIiiIHhB", int(r["bp_left"]), int(r["bp_right"]), int(r["indel_size"]), int(r["left"]), int(r["right"]),
                                int(r["name_id"]), int(r["nt_size"]), int(r["bp"]), int(r["unique"])))
        f.write(struct.pack("<i", len(box_off) - 1))
        for b in range(len(box_off) - 1):
            m = idx[int(box_off[b]):int(box_off[b + 1])]
            f.write(struct.pack("<i", m.size))
            f.write(m.astype("<u4").tobytes())
        f.write(struct.pack("<q", len(bam_bytes)))
        f.write(bam_bytes)


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


def build_emul_exe(src, tmp_path, name):
    """the same program linked against the CPU emulation of the ABI (tests/emul/abi_emul.cpp) instead of libpindel_pg.so"""
    exe = str(tmp_path / name)
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++11", "-Wall", "-Wextra", "-o", exe, src, "-L" + EMUL_DIR, "-lpg_abi_emul",
                           "-Wl,-rpath," + EMUL_DIR])
    return exe


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle(tmp_path):
    mirror_case(build_exe(), tmp_path)


def mirror_case(exe, tmp_path):
    ref = synth.make_reference(500_000, 73, n_runs=[(200_000, 800)])
    padded = synth.pad_reference(ref)
    batch = synth.make_reads(ref, 3000, 100, seed=74)
    write_input(tmp_path / "in.bin", [padded], batch, np.zeros(batch.n, np.int32))
    p = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    got = refhook.parse_stream(np.fromfile(tmp_path / "out.bin", dtype=np.int32), batch.n)
    exp = H.oracle_results([padded], batch, threads=4)
    H.compare(exp, got, "C++ host mirror")


def test_stage_mirrors_compile_and_fail_loudly_without_a_device(tmp_path):
    exe = build_exe(STAGES_SRC, STAGES_EXE)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("device present: covered by the gpu test")
    from oracle import boxsortoracle as B
    from pindel_b200 import bam, bgzf
    reads, box_off, idx = B.make_boxes(3, 1)
    write_stages_input(tmp_path / "in.bin", synth.make_discordant_pairs(3, 1), reads, box_off, idx, bgzf.compress(bam.synthetic_stream(20, 1)))
    p = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert p.returncode == 1 and "no CUDA device" in p.stderr


@pytest.mark.gpu
def test_cpp_stage_mirrors_match_oracles(tmp_path):
    """ModifyRP / Summarize, SortBoxes and InflateBam / IndexBam of the C++ mirror against the oracles of their stages."""
    stages_case(build_exe(STAGES_SRC, STAGES_EXE), tmp_path)


def stages_case(exe, tmp_path):
    from oracle import boxsortoracle as B, rporacle
    from pindel_b200 import bam, bgzf
    rows = synth.make_discordant_pairs(150, 41, span=3_000_000, background=200)
    reads, box_off, idx = B.make_boxes(300, 42)
    stream = bam.synthetic_stream(5000, 43)
    write_stages_input(tmp_path / "in.bin", rows, reads, box_off, idx, bgzf.compress(stream))
    p = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    o = np.fromfile(tmp_path / "out.bin", dtype=np.int32)
    t = o[:9 * rows.size].reshape(-1, 9).astype(np.int64)
    t[:, 1:5] &= 0xFFFFFFFF
    srt = rows[t[:, 0]]                                   # the order the mirror's std::sort(Compare2RP) left
    assert np.all(np.diff(srt["pos_a"].astype(np.int64)) <= 0)
    exp = rporacle.cluster(srt)
    for col, key in ((1, "pos_a"), (2, "pos_a1"), (3, "pos_b"), (4, "pos_b1"), (5, "visited"), (6, "report"), (7, "n_identical")):
        assert np.array_equal(t[:, col], exp[key].astype(np.int64)), key
    n_tags = np.bincount(exp["merged_into"], minlength=rows.size)       # a row holds its own tag and those merged into it
    assert np.array_equal(t[:, 8], np.where(exp["merged_into"] == np.arange(rows.size), n_tags, 0))
    p0 = 9 * rows.size
    e_idx, e_unique = B.sort_boxes(reads, box_off, idx)
    assert np.array_equal(o[p0:p0 + idx.size].astype(np.uint32), e_idx)
    assert np.array_equal(o[p0 + idx.size:p0 + idx.size + reads.size].astype(np.uint8), e_unique)
    q = p0 + idx.size + reads.size
    table = bam.serial_index(stream)
    assert o[q] == len(stream) and o[q + 1] == 1 and o[q + 2] == table.size
    rec = o[q + 3:q + 3 + 3 * table.size].reshape(-1, 3)
    assert np.array_equal(rec[:, 0].astype(np.int64), table["off"].astype(np.int64)) and np.array_equal(rec[:, 1], table["pos"]) and \
        np.array_equal(rec[:, 2], table["end"])
    # the inversion reporter's loops through SortBoxes(PG_BOX_INV_NT, ...): same boxes, keys as the C++ side derives them
    k = reads.copy()
    k["left"] = reads["left"]
    k["matched_d"] = np.where(reads["name_id"] & 1, ord("+"), ord("-"))
    k["read_length"] = np.where(reads["name_id"] % 3 == 0, 101, 100)
    i_idx, i_unique, i_event = B.sort_boxes_kind(4, k, box_off, idx)
    z = q + 3 + 3 * table.size
    assert np.array_equal(o[z:z + idx.size].astype(np.uint32), i_idx)
    assert np.array_equal(o[z + idx.size:z + idx.size + reads.size].astype(np.uint8), i_unique)
    assert np.array_equal(o[z + idx.size + reads.size:z + 2 * idx.size + reads.size].astype(np.uint8), i_event)


BD_SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_bd_test.cpp")
BD_EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_bd_test")


def bd_mirror_case(exe, tmp_path, seed):
    """Searcher::loadRegion + getCorrespondingSearchWindowClusters against the oracle's position sweep"""
    from oracle import bdoracle
    from tests import bd_helpers as BH
    chr_id, window, ev, q = BH.case(seed, seed % 2 == 0, n_queries=500)
    ev_sorted = np.array(bdoracle.sort_on_first(bdoracle.normalise(ev), BH.NAMES), dtype=np.int64).reshape(-1, 6)
    with open(tmp_path / "bd_in.bin", "wb") as f:
        f.write(struct.pack("<i", len(BH.NAMES)))
        for name, length in zip(BH.NAMES, BH.LENS):
            f.write(struct.pack("<i", len(name))); f.write(name.encode()); f.write(struct.pack("<I", length + 2 * BH.SPACER))
        f.write(struct.pack("<i", ev_sorted.shape[0]))
        for e in ev_sorted:
            f.write(struct.pack("<iIIiII", *[int(v) for v in e]))
        f.write(struct.pack("<iII", chr_id, window[0], window[1]))
        f.write(struct.pack("<i", q.size)); f.write(q.astype("<u4").tobytes())
    p = subprocess.run([exe, str(tmp_path / "bd_in.bin"), str(tmp_path / "bd_out.bin")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    o = np.fromfile(tmp_path / "bd_out.bin", dtype=np.int32)
    exp = bdoracle.clusters(ev, BH.NAMES, chr_id, window[0], window[1], q)
    pos = 0
    for k, e in enumerate(exp):
        c = int(o[pos]); pos += 1
        got = [(int(o[pos + 3 * j]), int(o[pos + 3 * j + 1]) & 0xFFFFFFFF, int(o[pos + 3 * j + 2]) & 0xFFFFFFFF) for j in range(c)]
        pos += 3 * c
        assert got == (e or []), (seed, k, int(q[k]), e, got)
    assert pos == o.size


@pytest.mark.skipif(not os.path.exists(os.path.join(EMUL_DIR, "libpg_abi_emul.so")), reason="oracle/_ref/libpg_abi_emul.so not built")
def test_search_mirror_host_logic_on_the_emulated_abi(tmp_path):
    """Searcher::GetCloseEnd / SearchFarEnds on SPLIT_READ vectors (marshalling into the SoA batch, runs back into UniquePoints,
    CleanUniquePoints, the flipped reads) without a GPU: same program, same 3000 reads, same comparison as the GPU test"""
    mirror_case(build_emul_exe(SRC, tmp_path, "host_mirror_test_emul"), tmp_path)


@pytest.mark.skipif(not os.path.exists(os.path.join(EMUL_DIR, "libpg_abi_emul.so")), reason="oracle/_ref/libpg_abi_emul.so not built")
def test_stage_mirrors_host_logic_on_the_emulated_abi(tmp_path):
    stages_case(build_emul_exe(STAGES_SRC, tmp_path, "host_mirror_stages_test_emul"), tmp_path)


@pytest.mark.skipif(not os.path.exists(os.path.join(EMUL_DIR, "libpg_abi_emul.so")), reason="oracle/_ref/libpg_abi_emul.so not built")
def test_bd_mirror_host_logic_on_the_emulated_abi(tmp_path):
    """the header's own work (event marshalling, name ranks, two-call buffer protocol) checked without a GPU: the same program
    linked against the CPU emulation of the ABI (tests/emul/abi_emul.cpp) instead of libpindel_pg.so"""
    exe = str(tmp_path / "host_mirror_bd_test_emul")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++11", "-Wall", "-Wextra", "-o", exe, BD_SRC, "-L" + EMUL_DIR, "-lpg_abi_emul",
                           "-Wl,-rpath," + EMUL_DIR])
    for seed in (1, 2, 3, 6):
        bd_mirror_case(exe, tmp_path, seed)
