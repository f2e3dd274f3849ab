"""scripts/bam_corruption_campaign.py -- CPU-only, meant to run under AddressSanitizer (scripts/sanitize_bam_corruption.sh): the
device programs that parse BAM records -- the record index (pg_bam.cuh: guesses, walks, CIGAR span), the discordant-pair
discovery and the split-read filter (pg_sr.cuh: names, flags, CIGARs, packed sequences, the aux-tag walker) -- on streams
whose records are DAMAGED but whose block_size chain is intact, so that the index accepts the stream and every later program
sees the damaged fields: l_read_name, n_cigar_op, l_seq, flags, positions, CIGAR operations, tag types and lengths overwritten at
random.  There is no reference answer for such input (htslib gives garbage or an error); what must hold is what the GPU
cannot show: no read outside the stream, the record table or a record (ASan stops the run), termination, and the programs'
own error reporting instead of a crash.      python scripts/bam_corruption_campaign.py [--rounds N] [--seed S]"""
import argparse
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def damage(rng, stream, first):
    b = bytearray(stream)
    # record starts by the serial walk of the intact stream
    starts, p = [], first
    while p + 4 <= len(b):
        bs = struct.unpack_from("<I", b, p)[0]
        starts.append((p, bs))
        p += 4 + bs
    for _ in range(int(rng.integers(1, 40))):
        p, bs = starts[int(rng.integers(0, len(starts)))]
        kind = int(rng.integers(0, 7))
        if kind == 0:
            b[p + 4 + 8] = int(rng.integers(0, 256))                                          # l_read_name
        elif kind == 1:
            struct.pack_into("<H", b, p + 4 + 12, int(rng.choice([0, 1, 3, 200, 65535])))     # n_cigar_op
        elif kind == 2:
            struct.pack_into("<I", b, p + 4 + 16, int(rng.choice([0, 1, 5000, 1 << 27, 0xFFFFFFFF])))   # l_seq
        elif kind == 3:
            struct.pack_into("<H", b, p + 4 + 14, int(rng.integers(0, 65536)))                # flag
        elif kind == 4:
            struct.pack_into("<i", b, p + 4 + int(rng.choice([0, 4, 20, 24])), int(rng.choice([-5, -1, 0, 7, 1 << 30, -(1 << 31)])))   # refID / pos / mate
        else:                                                                                # anywhere in the body (CIGAR, sequence, tags)
            for _ in range(int(rng.integers(1, 6))):
                b[p + 4 + int(rng.integers(32, max(33, bs)))] = int(rng.integers(0, 256))
    return bytes(b)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=300)
    ap.add_argument("--seed", type=int, default=1)
    run(ap.parse_args())


def run(a):
    from pindel_b200 import bam
    from tests import sr_fuzz_helpers as F
    from tests.emul import bam_emul, sr_emul
    blib, slib = bam_emul.build(), sr_emul.build()
    rng = np.random.default_rng(a.seed)
    indexed = refused = cands = rows = flagged = 0
    for r in range(a.rounds):
        stream = F.fuzz_stream(a.seed * 100_000 + r, n_names=int(rng.integers(20, 400)))
        first, _ = bam.first_record_offset(stream)
        bad = damage(rng, stream, first)
        try:
            recs, stats = bam_emul.index(blib, bad, int(rng.choice([0, 700, 4000, 65536])))
        except ValueError:
            refused += 1
            continue
        indexed += 1
        for tid, beg, end, isz, maq in F.QUERIES:
            out = bam_emul.rp_discovery(blib, bad, recs, tid, beg, end, isz, maq)
            rows += out[0].size + out[2].size
            q = sr_emul.query(tid, beg, end, isz, F.REFS[tid][1], min_anchor_quality=maq)
            try:
                c = sr_emul.candidates(slib, bad, recs, q, shuffle_seed=int(rng.integers(0, 1 << 30)))
                cands += c[0].size
                flagged += c[4] != 0
            except ValueError:
                flagged += 1
    print("campaign clean: %d damaged streams (%d indexed, %d refused by the index), %d pair rows and %d split-read candidates produced, "
          "%d filter calls ended with the program's own error flag; no crash%s" %
          (a.rounds, indexed, refused, rows, cands, flagged, ", no sanitizer report" if "asan" in os.environ.get("LD_PRELOAD", "") else ""))
    return indexed, rows, cands, flagged


if __name__ == "__main__":
    main()
