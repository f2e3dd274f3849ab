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
    # ... and the boxes of the DI / inversion reporters (inlined loops: the reporters themselves are bound) by pg_box_sort_kind
    m = re.search(r"(\d+) boxes of the DI / inversion reporters ordered and cut into events there", p.stderr)
    assert m and (int(m.group(1)) == 0) == bool(env.get("PG_BRIDGE_HOST_SORT")), p.stderr[-500:]
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
    # rest of row N3: the split-read candidates came from the device filter instead of the reader's per-record loop, and were
    # searched where the filter left them (no upload of their sequences) unless the scenario keeps the two steps apart
    m = re.search(r"(\d+) split-read candidates from the device filter, (\d+) of them searched where the filter left them", p.stderr)
    on = not (env.get("PG_BRIDGE_HOST_SR") or env.get("PG_BRIDGE_ZLIB"))
    assert m and (int(m.group(1)) > 0) == on, p.stderr[-300:]
    assert int(m.group(2)) == (0 if (env.get("PG_BRIDGE_SR_UNCHAINED") or not on) else int(m.group(1))), p.stderr[-300:]
    # rest of row N2: BDData::loadRegion handed the window's events to the library and SearchFarEnds sent no clusters --
    # pg_far_end looked every read's cluster up on the device (PG_BRIDGE_HOST_BD=1: the reference's mask on the host)
    m = re.search(r"(\d+) BreakDancer regions \((\d+) events\) loaded into the library, (\d+) cluster lookups there", p.stderr)
    assert m, p.stderr[-300:]
    if env.get("PG_BRIDGE_HOST_BD"):
        assert int(m.group(1)) == 0 and int(m.group(3)) == 0, m.group(0)
    else:
        assert int(m.group(1)) > 0 and int(m.group(2)) > 0 and int(m.group(3)) > 0, m.group(0)
    return p


def dense_events_live(demo, tmp_path, env):
    """Input dense in events of every reporter (tests/dense_events.py) -- above all deletions with non-template bases and
    inversions, whose reporters carry ordering, duplicate scan and event cut as inlined loops (SortOutputDI,
    OutputSorter::DoSortAndOutputInversions) -- so that all five supplied reporters run on crowded boxes.  No golden files:
    the pure reference binary (oracle/_ref/pindel_ref) runs beside the drop-in binary on the same seeded input."""
    from oracle import refhook
    from tests import dense_events
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "pindel_ref")
    ref, text, n = dense_events.make()
    fa = str(tmp_path / "dense.fa")
    refhook.write_fasta(fa, [("dense", ref)])
    (tmp_path / "dense.fa.fai").write_text("dense\t%d\t7\t60\t61\n" % ref.size)
    (tmp_path / "reads.txt").write_text(text)
    args = ["-f", fa, "-p", str(tmp_path / "reads.txt"), "-k", "-s", "-l", "-T", "1"]
    r = subprocess.run([ref_bin] + args + ["-o", str(tmp_path / "ref")], capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    p = _run(demo, args + ["-o", str(tmp_path / "dropin")], env)
    for suf in ("D", "SI", "INV", "TD", "LI", "BP", "CloseEndMapped"):
        a, b = (tmp_path / ("ref_" + suf)).read_bytes(), (tmp_path / ("dropin_" + suf)).read_bytes()
        assert a == b, "%s differs from the reference binary's file (%d vs %d bytes)" % (suf, len(b), len(a))
    n_events = {suf: sum(1 for line in (tmp_path / ("ref_" + suf)).read_text().split("\n") if line.startswith("####")) for suf in ("D", "SI", "INV", "TD")}
    n_di = len(re.findall(r"\tNT [1-9]", (tmp_path / "ref_D").read_text()))
    assert n_di >= 20 and n_events["INV"] >= 15 and n_events["SI"] >= 8 and n_events["TD"] >= 5 and n_events["D"] - n_di >= 8, (n_di, n_events)
    m = re.search(r"(\d+) boxes of the DI / inversion reporters", p.stderr)
    assert env.get("PG_BRIDGE_HOST_SORT") or int(m.group(1)) >= 40, m.group(0)
    return p


def interchr_live(demo, tmp_path, env):
    """A scenario none of the committed fixtures has: translocation-like clusters of pairs between two chromosomes with
    `-I true` (inter-chromosomal events are off by default, src/fn_parameters.cpp:163-166).  No golden files: the inputs are
    generated here from seeds and the pure reference binary (oracle/_ref/pindel_ref, which travels with the snapshot) runs
    beside the drop-in binary.  Exercises the inter-chromosomal list of the pair discovery and ModifyRP_InterChr /
    Summarize_InterChr through the bridge."""
    import struct
    import numpy as np
    from oracle import refhook
    from pindel_b200 import bam, bgzf
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "pindel_ref")
    rng = np.random.default_rng(7)
    refs = (("chrA", 300_000), ("chrB", 250_000))
    seqs = [rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), l) for _, l in refs]
    fa = str(tmp_path / "g.fa")
    refhook.write_fasta(fa, [(n, s) for (n, _), s in zip(refs, seqs)])
    off, fai = 0, []
    for n, l in refs:                                          # samtools faidx layout: name, length, offset, 60 bases / 61 bytes per line
        off += len(n) + 2
        fai.append("%s\t%d\t%d\t60\t61\n" % (n, l, off))
        off += l + (l + 59) // 60
    (tmp_path / "g.fa.fai").write_text("".join(fai))
    recs, k = [], 0
    code = {65: 1, 67: 2, 71: 4, 84: 8}

    def add(tid, pos, flag, mtid, mpos, tlen):
        rl = 100
        c = np.array([code[int(x)] for x in seqs[tid][pos:pos + rl]], dtype=np.uint8)
        packed = (c[0::2] << 4 | c[1::2]).astype(np.uint8).tobytes()
        nm = b"pair_%d\0" % k
        body = struct.pack("<iiBBHHHiiii", tid, pos, len(nm), 60, 4680, 1, flag, rl, mtid, mpos, tlen) + nm + struct.pack("<I", rl << 4) + packed + \
            bytes([30] * rl) + b"NMC\x00"
        recs.append((tid, pos, struct.pack("<I", len(body)) + body))

    for ev in range(12):
        pa, pb = int(rng.integers(20_000, 280_000)), int(rng.integers(20_000, 230_000))
        sa, sb = [(0, 0), (16, 0), (0, 16), (16, 16)][ev % 4]
        for _ in range(int(rng.integers(15, 40))):
            xa, xb = pa + int(rng.integers(-150, 150)), pb + int(rng.integers(-150, 150))
            for _ in range(int(rng.integers(1, 4))):          # duplicates: Summarize_InterChr wants identical rows
                add(0, xa, 1 | 64 | sa | (32 if sb else 0), 1, xb, 0)
                add(1, xb, 1 | 128 | sb | (32 if sa else 0), 0, xa, 0)
                k += 1
    for _ in range(3000):                                     # ordinary pairs around them
        tid = int(rng.integers(0, 2))
        p = int(rng.integers(1000, refs[tid][1] - 2000))
        add(tid, p, 1 | 64 | 32 | 2, tid, p + 400, 500)
        add(tid, p + 400, 1 | 128 | 16 | 2, tid, p, -500)
        k += 1
    recs.sort(key=lambda r: (r[0], r[1]))
    (tmp_path / "x.bam").write_bytes(bgzf.compress(bam.header(list(refs)) + b"".join(r[2] for r in recs)))
    (tmp_path / "x.bam.bai").write_bytes(b"BAI\x01\x00\x00\x00\x00")
    (tmp_path / "conf.txt").write_text("%s\t500\tS1\n" % (tmp_path / "x.bam"))
    args = ["-f", fa, "-i", str(tmp_path / "conf.txt"), "-c", "ALL", "-T", "1", "-I", "true"]
    r = subprocess.run([ref_bin] + args + ["-o", str(tmp_path / "ref")], capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    assert r.stdout.count("adding interchr") >= 10            # the scenario does produce inter-chromosomal events
    p = subprocess.run([demo] + args + ["-o", str(tmp_path / "got")], capture_output=True, text=True, timeout=1800, env=env)
    assert p.returncode == 0, p.stdout[-1000:] + p.stderr[-1000:]
    m = re.search(r"(\d+) read pairs clustered on the device \((\d+) launches\)", p.stderr)
    assert m and int(m.group(1)) > 1000 and int(m.group(2)) >= 4, p.stderr[-400:]
    for suf in ("D", "SI", "INV", "TD", "LI", "BP", "RP", "INT_final"):
        assert (tmp_path / ("got_" + suf)).read_bytes() == (tmp_path / ("ref_" + suf)).read_bytes(), "_%s differs from the reference binary's file" % suf
    assert (tmp_path / "ref_RP").read_bytes().count(b"\n") >= 20
    return p
