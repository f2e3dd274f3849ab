"""Drop-in check at the file level: the UNMODIFIED reference host program (command line, Pindel-text input, window loop,
SV drivers, reporter) with its two search seam functions served by libpindel_pg.so (oracle/ref_gpu_bridge.cpp, bound by
the dynamic linker -- no reference source is patched) must write byte-identical _D/_SI/_INV/_TD (and _LI/_BP/
_CloseEndMapped) files to the ones the pure reference binary wrote for the same inputs (tests/golden/colo_slice.*).
Needs oracle/_ref/pindel_gpu_demo, which is built in the container where /root/reference exists and travels to the
GPU box with the snapshot."""
import gzip
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "oracle", "_ref", "pindel_gpu_demo")
GOLD = os.path.join(ROOT, "tests", "golden")

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(DEMO), reason="oracle/_ref/pindel_gpu_demo not built (needs /root/reference at build time)")
def test_reference_host_with_gpu_search_writes_identical_files(tmp_path):
    fa = tmp_path / "colo_slice.fa"
    fa.write_bytes(gzip.open(os.path.join(GOLD, "colo_slice.fa.gz")).read())
    (tmp_path / "colo_slice.fa.fai").write_text("20s\t500000\t5\t60\t61\n")
    rd = tmp_path / "reads.txt"
    rd.write_bytes(gzip.open(os.path.join(GOLD, "colo_slice_reads.txt.gz")).read())
    # the freshly built product library, not a stale copy next to the demo
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "pindel_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    p = subprocess.run([DEMO, "-f", str(fa), "-p", str(rd), "-o", str(tmp_path / "colo_slice.out"), "-k", "-s", "-l", "-T", "1"],
                       capture_output=True, text=True, timeout=1200, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "served by libpindel_pg" in p.stderr and " 0 GetCloseEnd calls" not in p.stderr
    assert "classified by the SV drivers on the device" in p.stderr and "; 0 reads classified" not in p.stderr
    for suf in ("D", "SI", "INV", "TD", "LI", "BP", "CloseEndMapped"):
        got = (tmp_path / ("colo_slice.out_" + suf)).read_bytes()
        exp = gzip.open(os.path.join(GOLD, "colo_slice.out_%s.gz" % suf)).read()
        assert got == exp, "_%s differs from the reference's file (%d vs %d bytes)" % (suf, len(got), len(exp))
    # the same through the per-read GetCloseEnd shim (one-read batches): the seam function itself, unbatched
    env2 = dict(env, PG_BRIDGE_UNBATCHED="1")
    p2 = subprocess.run([DEMO, "-f", str(fa), "-p", str(rd), "-o", str(tmp_path / "unbatched.out"), "-k", "-s", "-l", "-T", "1"],
                        capture_output=True, text=True, timeout=1200, env=env2)
    assert p2.returncode == 0, p2.stdout[-2000:] + p2.stderr[-2000:]
    for suf in ("D", "SI", "INV", "TD"):
        got = (tmp_path / ("unbatched.out_" + suf)).read_bytes()
        exp = gzip.open(os.path.join(GOLD, "colo_slice.out_%s.gz" % suf)).read()
        assert got == exp, "unbatched _%s differs" % suf


@pytest.mark.skipif(not os.path.exists(DEMO), reason="oracle/_ref/pindel_gpu_demo not built (needs /root/reference at build time)")
def test_reference_host_bam_input_with_gpu_search(tmp_path):
    """BAM-config input (-i): the reference's BAM reader, ReadBuffer::flush (here: the batched binding), discordant-pair
    discovery and BreakDancer-like clusters feeding multi-region far-end searches; golden files written by the pure
    reference binary for the same inputs (tests/golden/sim_prefix.*)."""
    import shutil
    fa = tmp_path / "sim_prefix.fa"
    fa.write_bytes(gzip.open(os.path.join(GOLD, "sim_prefix.fa.gz")).read())
    shutil.copy(os.path.join(GOLD, "sim_prefix.fa.fai"), tmp_path / "sim_prefix.fa.fai")
    shutil.copy(os.path.join(GOLD, "sim_prefix.bam"), tmp_path / "sim_prefix.bam")
    (tmp_path / "sim_prefix.bam.bai").write_bytes(b"BAI\x01\x00\x00\x00\x00")
    (tmp_path / "conf.txt").write_text("%s\t500\tSIM1CHRVS2\n" % (tmp_path / "sim_prefix.bam"))
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "pindel_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    p = subprocess.run([DEMO, "-f", str(fa), "-i", str(tmp_path / "conf.txt"), "-o", str(tmp_path / "sim_prefix.out"), "-k", "-s", "-l", "-T", "1"],
                       capture_output=True, text=True, timeout=1200, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "served by libpindel_pg" in p.stderr and " 0 GetCloseEnd calls" not in p.stderr
    assert "classified by the SV drivers on the device" in p.stderr and "; 0 reads classified" not in p.stderr
    for suf in ("D", "SI", "INV", "TD", "LI", "BP", "RP", "CloseEndMapped"):
        got = (tmp_path / ("sim_prefix.out_" + suf)).read_bytes()
        exp = gzip.open(os.path.join(GOLD, "sim_prefix.out_%s.gz" % suf)).read()
        assert got == exp, "_%s differs from the reference's file (%d vs %d bytes)" % (suf, len(got), len(exp))
