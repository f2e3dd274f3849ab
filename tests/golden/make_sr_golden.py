"""Golden vectors of the reader's split-read filter (rest of row N3: ReadInBamReads_SR, src/reader.cpp:483-560, fetch_func_SR
:1099-1151, build_record_SR :799-900, build_record_RefRead :903-923) made by the REFERENCE ITSELF through
oracle/ref_hook.cpp pgref_sr_candidates, on the committed BAM fixture and on a seeded synthetic BAM of named mate pairs.
Run in the container where /root/reference exists (one process per fasta):
    python -m tests.golden.make_sr_golden
The .npz holds, per query, the candidate reads the reference handed to its ReadBuffer ([k, 7]: FNV-1a of the name, MatchedD,
MatchedRelPos, MS, InsertSize, read length, FNV-1a of UnmatchedSeq) and the RefSupportingReads it kept ([m, 3]: Pos, MQ,
ReadLength)."""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests.golden import make_rpd_golden as G      # noqa: E402

SYNTH_REFS = (("chr1", 400_000), ("chr2", 300_000))
SYNTH = dict(n_pairs=3000, seed=31)
# case -> list of (chr, start, end, insert_size, -A minimal anchor quality)
QUERIES = {
    "sim_prefix": [(0, 0, 10_000_000, 500, 0), (0, 100_000, 100_400, 500, 0), (0, 110_000, 120_000, 250, 30), (0, 390_000, 400_000, 500, 0)],
    "synthetic": [(0, 0, 400_000, 500, 0), (1, 0, 300_000, 400, 0), (0, 0, 400_000, 500, 30), (1, 100_000, 300_000, 400, 11), (0, 20_000, 30_000, 400, 0),
                  (0, 1000, 1100, 500, 0)],
}

_NIB = {c: i for i, c in enumerate(b"=ACMGRSVTWYHKDBN")}


def pack_seq(seq: bytes) -> bytes:
    nib = [_NIB[c] for c in seq] + [0]
    return bytes((nib[2 * i] << 4) | nib[2 * i + 1] for i in range((len(seq) + 1) // 2))


def paired_stream(n_pairs: int, seed: int, refs=SYNTH_REFS) -> bytes:
    """A coordinate-sorted BAM stream of named mate pairs, with what the filter branches on: unmapped mates placed at their
    anchor, clipped / gapped / mismatching mates, N runs at either end and inside, short reads, reverse strands, secondary
    and duplicate flags, low mapping qualities, names that occur once or three times."""
    import struct
    from pindel_b200 import bam
    rng = np.random.default_rng(seed)
    recs = []

    def rec(tid, pos, name, flag, mapq, cig, seq, mtid, mpos, tlen, aux):
        nm = name + b"\0"
        c = b"".join(struct.pack("<I", (ln << 4) | op) for op, ln in cig)
        body = struct.pack("<iiBBHHHiiii", tid, pos, len(nm), mapq, 4680, len(cig), flag, len(seq), mtid, mpos, tlen) + nm + c + \
            pack_seq(seq) + bytes(rng.integers(0, 42, len(seq), dtype=np.uint8)) + aux
        recs.append((tid, pos, len(recs), struct.pack("<I", len(body)) + body))

    def sequence(rl):
        s = bytearray(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), rl).tobytes())
        k = rng.random()
        if k < 0.15:
            a = int(rng.integers(1, 8)); s[:a] = b"N" * a
        elif k < 0.3:
            a = int(rng.integers(1, 8)); s[rl - a:] = b"N" * a
        elif k < 0.4:
            for p in rng.integers(0, rl, int(rng.integers(1, rl // 5))):
                s[p] = ord("N")
        elif k < 0.43:
            s[:] = b"N" * rl
        elif k < 0.5:          # IUPAC codes and '=' at the ends and inside: NUL after a reverse complement, stripped when trailing
            a, b = int(rng.integers(0, 4)), int(rng.integers(0, 4))
            s[:a] = rng.choice(np.frombuffer(b"=MRSVWYHKDB", dtype=np.uint8), a).tobytes()
            s[rl - b:] = rng.choice(np.frombuffer(b"=MRSVWYHKDB", dtype=np.uint8), b).tobytes()
            s[rl // 2] = ord("R")
        return bytes(s)

    def cigar(rl):
        k = rng.random()
        if k < 0.55:
            return [(0, rl)]
        if k < 0.7:
            a = int(rng.integers(5, rl - 5)); return [(4, a), (0, rl - a)]
        if k < 0.8:
            a = int(rng.integers(5, rl - 5)); return [(0, rl - a), (4, a)]
        if k < 0.88:
            a = int(rng.integers(5, rl - 10)); return [(0, a), (2, int(rng.integers(1, 40))), (0, rl - a)]
        if k < 0.94:
            a = int(rng.integers(5, rl - 10)); return [(0, a), (1, 4), (0, rl - a - 4)]
        return [(5, 10), (0, rl)]

    def aux():
        k = rng.random()
        out = b""
        if k < 0.75:
            out += b"NMC" + bytes([int(rng.choice([0, 0, 0, 1, 2, 3, 6]))])
        return out + b"RGZgrp%d\0" % int(rng.integers(0, 3))

    for i in range(n_pairs):
        tid = int(rng.integers(0, len(refs)))
        length = refs[tid][1]
        rl1, rl2 = (int(rng.choice([18, 36, 76, 100, 150])) for _ in range(2))
        p1 = int(rng.integers(0, length - 2000))
        p2 = p1 + int(rng.integers(0, 900))
        name = b"pair_%d" % i
        q1, q2 = (int(rng.choice([0, 10, 29, 30, 60])) for _ in range(2))
        r1 = 16 if rng.random() < 0.5 else 0
        r2 = 16 if rng.random() < 0.5 else 0
        extra1 = int(rng.choice([0, 0, 0, 0, 0x100, 0x200, 0x400]))
        extra2 = int(rng.choice([0, 0, 0, 0, 0x100, 0x200, 0x400]))
        k = rng.random()
        if k < 0.2:            # mate 2 unmapped, stored at the anchor
            rec(tid, p1, name, 1 | 0x40 | 8 | r1 | extra1, q1, cigar(rl1), sequence(rl1), tid, p1, 0, aux())
            rec(tid, p1, name, 1 | 0x80 | 4 | (32 if r1 else 0), 0, [], sequence(rl2), tid, p1, 0, b"")
        elif k < 0.25:         # a name that occurs once in the window
            rec(tid, p1, name, 1 | 0x40 | r1 | extra1, q1, cigar(rl1), sequence(rl1), 1 - tid if len(refs) > 1 else tid, p2, 0, aux())
        elif k < 0.3:          # a name three times: the third starts a new wait
            rec(tid, p1, name, 1 | 0x40 | r1, q1, cigar(rl1), sequence(rl1), tid, p2, p2 - p1, aux())
            rec(tid, p2, name, 1 | 0x80 | r2, q2, cigar(rl2), sequence(rl2), tid, p1, p1 - p2, aux())
            rec(tid, p2 + 40, name, 1 | 0x80 | 0x800 | r2, q2, cigar(rl2), sequence(rl2), tid, p1, p1 - p2, aux())
        else:
            rec(tid, p1, name, 1 | 0x40 | r1 | (32 if r2 else 0) | extra1, q1, cigar(rl1), sequence(rl1), tid, p2, p2 - p1 + rl2, aux())
            rec(tid, p2, name, 1 | 0x80 | r2 | (32 if r1 else 0) | extra2, q2, cigar(rl2), sequence(rl2), tid, p1, p1 - p2 - rl2, aux())
    recs.sort(key=lambda r: (r[0], r[1], r[2]))
    return bam.header(list(refs)) + b"".join(r[3] for r in recs)


def synthetic_bam() -> bytes:
    from pindel_b200 import bgzf
    return bgzf.compress(paired_stream(**SYNTH))


def stage(case: str, tmp: str):
    if case == "sim_prefix":
        return G.stage(case, tmp)
    from oracle import refhook
    rng = np.random.default_rng(1)
    fa = os.path.join(tmp, "synthetic.fa")
    refhook.write_fasta(fa, [(n, rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), l)) for n, l in SYNTH_REFS])
    bam_path = os.path.join(tmp, "synthetic.bam")
    open(bam_path, "wb").write(synthetic_bam())
    open(bam_path + ".bai", "wb").write(b"BAI\x01\x00\x00\x00\x00")
    return fa, bam_path


def run_case(case: str):
    from oracle import refhook
    tmp = tempfile.mkdtemp()
    fa, bam_path = stage(case, tmp)
    h = refhook.RefHook(fa, [], workdir=tmp)
    out = {"chr_len": np.array([h.chr_seq(c).size - 200000 for c in range(h.n_chr)], dtype=np.int64)}
    for k, q in enumerate(QUERIES[case]):
        cand, refs = h.sr_candidates(bam_path, *q)
        out["cand_%d" % k], out["refs_%d" % k] = cand, refs
        print(case, q, cand.shape[0], "candidates,", refs.shape[0], "reference-supporting reads")
    np.savez_compressed(os.path.join(HERE, "sr_%s.npz" % case), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for case in QUERIES:       # the reference keeps process-global state: one process per genome
            subprocess.check_call([sys.executable, "-m", "tests.golden.make_sr_golden", case], cwd=ROOT)
