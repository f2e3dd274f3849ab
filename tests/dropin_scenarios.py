"""The two file-level drop-in scenarios (Pindel-text input, BAM-config input), parameterised by the host binary: the
reference host + bridge on libpindel_pg.so (GPU test) or on the CPU emulators of the device programs (CPU test of the
bridge's host logic).  Golden files were written by the pure reference binary (tests/golden/)."""
import gzip
import os
import re
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _run(demo, args, env):
    p = subprocess.run([demo] + args, capture_output=True, text=True, timeout=1800, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "served by libpindel_pg" in p.stderr and " 0 GetCloseEnd calls" not in p.stderr
    # every window is classified by ONE call (all seven drivers in one launch), and the drivers replayed its records
    assert "classified by the SV drivers on the device" in p.stderr and "; 0 reads classified" not in p.stderr
    # ... on the batch SearchFarEnds left on the device, unless the scenario switches that off
    m = re.search(r"; (\d+) reads classified .* (\d+) of the reads without a second upload", p.stderr)
    assert m, p.stderr[-500:]
    assert int(m.group(2)) == (0 if env.get("PG_BRIDGE_NO_RESIDENT") else int(m.group(1))), m.group(0)
    # row N4: the boxes the D / SI / TD reporters sort were ordered by pg_box_sort, one call per driver
    m = re.search(r"(\d+) boxes ordered on the device \((\d+) launches\), (\d+) by the reference", p.stderr)
    assert m, p.stderr[-500:]
    if env.get("PG_BRIDGE_HOST_SORT"):
        assert int(m.group(1)) == 0 and int(m.group(3)) > 0, m.group(0)
    else:
        assert int(m.group(1)) > 0 and int(m.group(3)) == 0 and int(m.group(2)) <= 8, m.group(0)
    return p


def _compare(tmp_path, prefix, gold_prefix, suffixes):
    for suf in suffixes:
        got = (tmp_path / ("%s_%s" % (prefix, suf))).read_bytes()
        exp = gzip.open(os.path.join(GOLD, "%s.out_%s.gz" % (gold_prefix, suf))).read()
        assert got == exp, "%s_%s differs from the reference's file (%d vs %d bytes)" % (prefix, suf, len(got), len(exp))


def colo_slice(demo, tmp_path, env, unbatched_too=True):
    fa = tmp_path / "colo_slice.fa"
    fa.write_bytes(gzip.open(os.path.join(GOLD, "colo_slice.fa.gz")).read())
    (tmp_path / "colo_slice.fa.fai").write_text("20s\t500000\t5\t60\t61\n")
    rd = tmp_path / "reads.txt"
    rd.write_bytes(gzip.open(os.path.join(GOLD, "colo_slice_reads.txt.gz")).read())
    p = _run(demo, ["-f", str(fa), "-p", str(rd), "-o", str(tmp_path / "colo_slice.out"), "-k", "-s", "-l", "-T", "1"], env)
    _compare(tmp_path, "colo_slice.out", "colo_slice", ("D", "SI", "INV", "TD", "LI", "BP", "CloseEndMapped"))
    if unbatched_too:   # the same through the per-read GetCloseEnd shim (one-read batches): the seam function itself
        _run(demo, ["-f", str(fa), "-p", str(rd), "-o", str(tmp_path / "unbatched.out"), "-k", "-s", "-l", "-T", "1"],
             dict(env, PG_BRIDGE_UNBATCHED="1"))
        _compare(tmp_path, "unbatched.out", "colo_slice", ("D", "SI", "INV", "TD"))
    return p


def sim_prefix_bam(demo, tmp_path, env):
    """BAM-config input (-i): the reference's BAM reader, ReadBuffer::flush (here: the batched binding), discordant-pair
    discovery and BreakDancer-like clusters feeding multi-region far-end searches."""
    fa = tmp_path / "sim_prefix.fa"
    fa.write_bytes(gzip.open(os.path.join(GOLD, "sim_prefix.fa.gz")).read())
    shutil.copy(os.path.join(GOLD, "sim_prefix.fa.fai"), tmp_path / "sim_prefix.fa.fai")
    shutil.copy(os.path.join(GOLD, "sim_prefix.bam"), tmp_path / "sim_prefix.bam")
    (tmp_path / "sim_prefix.bam.bai").write_bytes(b"BAI\x01\x00\x00\x00\x00")
    (tmp_path / "conf.txt").write_text("%s\t500\tSIM1CHRVS2\n" % (tmp_path / "sim_prefix.bam"))
    p = _run(demo, ["-f", str(fa), "-i", str(tmp_path / "conf.txt"), "-o", str(tmp_path / "sim_prefix.out"), "-k", "-s", "-l", "-T", "1"], env)
    _compare(tmp_path, "sim_prefix.out", "sim_prefix", ("D", "SI", "INV", "TD", "LI", "BP", "RP", "CloseEndMapped"))
    # row N2: the discordant pairs of the window went through pg_rp_modify / pg_rp_summarize (ModifyRP / Summarize bound by symbol)
    m = re.search(r"(\d+) read pairs clustered on the device \((\d+) launches\)", p.stderr)
    assert m and int(m.group(1)) > 0 and int(m.group(2)) > 0, p.stderr[-500:]
    # row N3 (first half): the BAM file was inflated by pg_bgzf_inflate
    m = re.search(r"(\d+) BAM bytes inflated and (\d+) records indexed on the device, (\d+) discordant pairs found there", p.stderr)
    size = len(gzip.decompress((tmp_path / "sim_prefix.bam").read_bytes()))      # the reference opens the file several times
    assert m and (int(m.group(1)) == 0 if env.get("PG_BRIDGE_ZLIB") else (int(m.group(1)) > 0 and int(m.group(1)) % size == 0)), p.stderr[-300:]
    assert (int(m.group(2)) == 0) == bool(env.get("PG_BRIDGE_ZLIB")), p.stderr[-300:]
    # row N2, first half: the discordant pairs came from pg_bam_rp_discovery (7928 in this fixture, as the reference finds)
    assert int(m.group(3)) == (0 if env.get("PG_BRIDGE_ZLIB") else 7928), p.stderr[-300:]
    return p
