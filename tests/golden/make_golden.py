"""tests/golden/make_golden.py -- regenerates every fixture in this directory FROM THE REFERENCE ITSELF.

Run in the build container (needs /root/reference and oracle/_ref built by `make -C oracle ref`):

    python tests/golden/make_golden.py

What it writes (all small; the inputs are either regenerable from a seed or stored next to the outputs):
  synth_<name>.npz      per-read UniquePoint streams dumped from the reference (oracle/_ref/libpindel_ref.so running
                        the untouched GetCloseEnd / CleanUniquePoints / SearchFarEnds) on seeded synthetic batches
  edge_cases.npz        hand-made reads (N / IUPAC / lower case / trailing junk / short / near chromosome ends) + the
                        reference's answers
  sim_prefix.*          a prefix of test/SmallTest/sim1chrVs2.bam (re-blocked as BGZF) + its reference: the BAM-input
                        scenario, with the _D/_SI/_INV/_TD/... files written by the reference binary
  colo_slice.*          a 500 kbp slice of test/SmallTest/hs_ref_chr20.fa (with the start of the centromere N run)
                        and the reads of test/SmallTest/COLO-829_20-p_ok-1.txt anchored in it: inputs, the reference's
                        per-read streams, and the _D/_SI/_INV/_TD files written by the reference binary
The reference keeps process-global state, so every configuration runs in its own subprocess.
"""
from __future__ import annotations

import gzip
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

SYNTH = {
    # name: (ref_bases, ref_seed, n_runs, n_reads, L, read_seed, extra pindel args, generator kwargs)
    "L100_default": (600_000, 101, [(300_000, 4000), (450_011, 3)], 1200, 100, 201, [], {}),
    "L150_default": (600_000, 102, [(200_000, 2500)], 1000, 150, 202, [], {}),
    "L75_x4_u05": (800_000, 103, [], 800, 75, 203, ["-x", "4", "-u", "0.05"], {"max_del": 20000}),
    "L100_a2_m5_H10_e03": (600_000, 104, [(100_000, 1000)], 800, 100, 204, ["-a", "2", "-m", "5", "-H", "10", "-e", "0.03"], {}),
    "L250_default": (600_000, 105, [], 400, 250, 205, [], {}),
}


def worker_synth(name):
    from oracle import refhook
    from pindel_b200 import synth
    ref_bases, ref_seed, n_runs, n, L, seed, extra, gkw = SYNTH[name]
    ref = synth.make_reference(ref_bases, ref_seed, n_runs=n_runs)
    batch = synth.make_reads(ref, n, L, seed=seed, **gkw)
    tmp = tempfile.mkdtemp()
    fa = os.path.join(tmp, "ref.fa")
    refhook.write_fasta(fa, [("chrS", ref)])
    h = refhook.RefHook(fa, extra, workdir=tmp)
    h.set_window(0, 0, ref.size)
    blob, off = batch.blob()
    stream, _ = h.search(blob, off, batch.matched_d, batch.matched_rel_pos, batch.insert_size, 0, True, 4)
    np.savez_compressed(os.path.join(HERE, "synth_%s.npz" % name), stream=stream, n=n,
                        chr_len=h.chr_seq(0).size, table=h.max_mismatch(), args=np.array(extra, dtype="U16"),
                        seq_digest=np.frombuffer(__import__("hashlib").sha256(batch.seq.tobytes()).digest(), dtype=np.uint8))


def edge_reads(ref):
    """Hand-made corner cases; positions refer to a 300 kbp random reference."""
    from pindel_b200 import synth
    comp = {65: 84, 67: 71, 71: 67, 84: 65, 78: 78}
    base = synth.make_reads(ref, 60, 100, seed=77, sub_rate=0.0, n_rate=0.0, decoy_frac=0.0)
    seqs, md, pos, isz = [], [], [], []

    def add(s, d, p, i=500):
        seqs.append(bytes(s)); md.append(d); pos.append(p); isz.append(i)

    for i in range(base.n):
        s = bytearray(base.seq[i].tobytes())
        k = i % 12
        if k == 1: s[0] = ord("N")                       # first char N
        if k == 2: s[-1] = ord("N")                      # last char N
        if k == 3: s[0] = ord("M")                       # IUPAC at the start: becomes a trailing NUL after a flip
        if k == 4: s[-1] = ord("R")
        if k == 5: s[10] = ord("a")                      # lower case never matches
        if k == 6: s += b"\r"                            # trailing junk is stripped by setUnmatchedSeq
        if k == 7: s[40:45] = b"NNNNN"
        if k == 8: s = s[:31]                            # shorter than one 32-base word
        if k == 9: s = s[:9]                             # shorter than bp_start + 2
        if k == 10: s[0:3] = b"YKM"
        if k == 11: s = bytearray(bytes(comp.get(c, 78) for c in reversed(s)))   # wrong strand: forces the flip
        add(s, int(base.matched_d[i]), int(base.matched_rel_pos[i]))
    # anchors at the very ends of the chromosome (windows reach into the N spacer) and a few odd insert sizes
    add(ref[20:120].tobytes(), ord("-"), 200)
    add(ref[100:200].tobytes(), ord("-"), 150, 900)
    add(bytes(comp[c] for c in reversed(ref[ref.size - 150:ref.size - 50].tobytes())), ord("+"), ref.size - 300)
    add(ref[ref.size - 100:].tobytes(), ord("-"), ref.size - 10)
    add(b"ACGT" * 25, ord("+"), 5000)
    add(b"A" * 100, ord("-"), 7000)
    add(b"N" * 100, ord("+"), 9000)
    add(ref[50_000:50_100].tobytes(), ord("-"), 50_300, 0)          # insert size 0: empty windows
    add(ref[60_000:60_100].tobytes(), ord("-"), 60_300, 30)         # tiny window
    add(ref[70_000:70_100].tobytes(), ord("*"), 70_300)             # MatchedD neither + nor -
    return seqs, np.array(md, np.uint8), np.array(pos, np.uint32), np.array(isz, np.int16)


def worker_edge():
    from oracle import refhook
    from pindel_b200 import synth
    ref = synth.make_reference(300_000, 106, n_runs=[(150_000, 700)])
    seqs, md, pos, isz = edge_reads(ref)
    off = np.zeros(len(seqs) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(s) for s in seqs])
    blob = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    tmp = tempfile.mkdtemp()
    fa = os.path.join(tmp, "ref.fa")
    refhook.write_fasta(fa, [("chrS", ref)])
    h = refhook.RefHook(fa, [], workdir=tmp)
    h.set_window(0, 0, ref.size)
    stream, _ = h.search(blob, off, md, pos, isz, 0, True, 1)
    np.savez_compressed(os.path.join(HERE, "edge_cases.npz"), stream=stream, blob=blob, off=off, md=md, pos=pos, isz=isz,
                        chr_len=h.chr_seq(0).size)


SLICE = (25_900_000, 26_400_000)


def worker_colo():
    from oracle import refhook
    src = "/root/reference/test/SmallTest/"
    txt = open(src + "hs_ref_chr20.fa", "rb").read()
    i = txt.index(b"\n")
    seq = np.frombuffer(txt[i + 1:].replace(b"\n", b""), dtype=np.uint8)
    lo, hi = SLICE
    sl = seq[lo:hi]
    names, seqs, md, pos, ms, isz, tags = [], [], [], [], [], [], []
    with open(src + "COLO-829_20-p_ok-1.txt") as f:
        while True:
            nm = f.readline()
            if not nm:
                break
            s = f.readline().rstrip("\n")
            l = f.readline().split()
            p = int(l[2])
            if lo + 2000 <= p < hi - 2000:
                names.append(nm.rstrip("\n")); seqs.append(s); md.append(l[0]); pos.append(p - lo); ms.append(l[3]); isz.append(l[4]); tags.append(l[5])
    with gzip.open(os.path.join(HERE, "colo_slice.fa.gz"), "wb", compresslevel=9) as g:
        g.write(b">20s\n")
        for k in range(0, sl.size, 60):
            g.write(sl[k:k + 60].tobytes() + b"\n")
    text = "".join("%s\n%s\n%s\t20s\t%d\t%s\t%s\t%s\n" % (names[k], seqs[k], md[k], pos[k], ms[k], isz[k], tags[k]) for k in range(len(seqs)))
    with gzip.open(os.path.join(HERE, "colo_slice_reads.txt.gz"), "wb", compresslevel=9) as g:
        g.write(text.encode())
    tmp = tempfile.mkdtemp()
    fa = os.path.join(tmp, "colo_slice.fa")
    with open(fa, "wb") as f:
        f.write(gzip.open(os.path.join(HERE, "colo_slice.fa.gz")).read())
    open(fa + ".fai", "w").write("20s\t%d\t5\t60\t61\n" % sl.size)
    rd = os.path.join(tmp, "reads.txt")
    open(rd, "w").write(text)
    # (1) whole-program goldens from the reference binary, same switches as devtools/run_regressiontests.sh:5
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "pindel_ref"), "-f", fa, "-p", rd, "-o", os.path.join(tmp, "colo_slice.out"),
                           "-k", "-s", "-l", "-T", "1"], stdout=subprocess.DEVNULL)
    for suf in ("D", "SI", "INV", "TD", "LI", "BP", "CloseEndMapped"):
        data = open(os.path.join(tmp, "colo_slice.out_" + suf), "rb").read()
        with gzip.open(os.path.join(HERE, "colo_slice.out_%s.gz" % suf), "wb", compresslevel=9) as g:
            g.write(data)
    # (2) per-read streams from the hook
    h = refhook.RefHook(fa, [], workdir=tmp)
    h.set_window(0, 0, sl.size)
    off = np.zeros(len(seqs) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(s) for s in seqs])
    blob = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
    stream, _ = h.search(blob, off, np.array([ord(c) for c in md], np.uint8), np.array(pos, np.uint32),
                         np.array([int(x) for x in isz], np.int16), 0, True, 4)
    np.savez_compressed(os.path.join(HERE, "colo_slice_stream.npz"), stream=stream, n=len(seqs), chr_len=h.chr_seq(0).size)
    print("colo_slice: %d reads" % len(seqs))


def bgzf_write(path, payload: bytes):
    """Valid BGZF: gzip members of <= 65280 input bytes with the BC extra field, then the empty EOF block."""
    import struct
    import zlib
    with open(path, "wb") as f:
        def block(chunk):
            co = zlib.compressobj(9, zlib.DEFLATED, -15)
            comp = co.compress(chunk) + co.flush()
            bsize = 12 + 6 + len(comp) + 8
            f.write(struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize - 1))
            f.write(comp)
            f.write(struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
        for k in range(0, len(payload), 65280):
            block(payload[k:k + 65280])
        block(b"")


BAM_CUT = (100_000, 135_000)   # keep the alignments that start in this range (inversion + tandem duplication + RP clusters)


def worker_bam():
    """BAM-input scenario (devtools/run_regressiontests.sh:7) on a prefix of test/SmallTest/sim1chrVs2.bam: exercises
    the reference's BAM reader (through oracle/shim/minihts.cpp), ReadBuffer::flush, discordant-pair discovery and the
    BreakDancer-like clusters that feed multi-region far-end searches."""
    import struct
    src = "/root/reference/test/SmallTest/"
    raw = gzip.open(src + "sim1chrVs2.bam").read()
    assert raw[:4] == b"BAM\1"
    p = 4
    l_text = struct.unpack_from("<i", raw, p)[0]; p += 4 + l_text
    n_ref = struct.unpack_from("<i", raw, p)[0]; p += 4
    for _ in range(n_ref):
        l = struct.unpack_from("<i", raw, p)[0]; p += 4 + l + 4
    out = bytearray(raw[:p])
    kept = 0
    while p + 4 <= len(raw):
        bs = struct.unpack_from("<i", raw, p)[0]
        pos = struct.unpack_from("<i", raw, p + 8)[0]
        if BAM_CUT[0] <= pos < BAM_CUT[1]:
            out += raw[p:p + 4 + bs]; kept += 1
        p += 4 + bs
    bgzf_write(os.path.join(HERE, "sim_prefix.bam"), bytes(out))
    fa = open(src + "sim1chrVs2.fa", "rb").read()
    with gzip.open(os.path.join(HERE, "sim_prefix.fa.gz"), "wb", compresslevel=9) as g:
        g.write(fa)
    fai = open(src + "sim1chrVs2.fa.fai").read()
    open(os.path.join(HERE, "sim_prefix.fa.fai"), "w").write(fai)
    tmp = tempfile.mkdtemp()
    open(os.path.join(tmp, "sim_prefix.fa"), "wb").write(fa)
    open(os.path.join(tmp, "sim_prefix.fa.fai"), "w").write(fai)
    import shutil
    shutil.copy(os.path.join(HERE, "sim_prefix.bam"), os.path.join(tmp, "sim_prefix.bam"))
    # the reference insists on an index file next to the BAM (src/pindel.cpp:690-697); the test reader ignores its content
    open(os.path.join(tmp, "sim_prefix.bam.bai"), "wb").write(b"BAI\x01\x00\x00\x00\x00")
    open(os.path.join(tmp, "conf.txt"), "w").write("%s\t500\tSIM1CHRVS2\n" % os.path.join(tmp, "sim_prefix.bam"))
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "pindel_ref"), "-f", os.path.join(tmp, "sim_prefix.fa"), "-i",
                           os.path.join(tmp, "conf.txt"), "-o", os.path.join(tmp, "sim_prefix.out"), "-k", "-s", "-l", "-T", "1"],
                          stdout=subprocess.DEVNULL)
    for suf in ("D", "SI", "INV", "TD", "LI", "BP", "RP", "CloseEndMapped"):
        data = open(os.path.join(tmp, "sim_prefix.out_" + suf), "rb").read()
        with gzip.open(os.path.join(HERE, "sim_prefix.out_%s.gz" % suf), "wb", compresslevel=9) as g:
            g.write(data)
    print("sim_prefix: %d alignments kept" % kept)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        kind, _, name = sys.argv[1].partition(":")
        {"synth": lambda: worker_synth(name), "edge": worker_edge, "colo": worker_colo, "bam": worker_bam}[kind]()
    else:
        for nm in SYNTH:
            subprocess.check_call([sys.executable, __file__, "synth:" + nm])
        subprocess.check_call([sys.executable, __file__, "edge"])
        subprocess.check_call([sys.executable, __file__, "colo"])
        subprocess.check_call([sys.executable, __file__, "bam"])
        subprocess.check_call(["ls", "-la", HERE])
