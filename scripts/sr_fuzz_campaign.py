"""scripts/sr_fuzz_campaign.py -- a longer three-way differential run of the reader's split-read filter: the REFERENCE ITSELF
(ReadInBamReads_SR through oracle/_ref/libpindel_ref.so), the Python oracle (oracle/sroracle.py) and the device functions' host
build (tests/emul/sr_emul.cpp, name table filled in shuffled order with keys cut to 12 bits), on random BAM streams that visit
the format's corners (tests/sr_fuzz_helpers.py).  Not part of the pytest suite; run it after touching pg_sr.cuh:
    python scripts/sr_fuzz_campaign.py [--rounds N]"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=40)
    ap.add_argument("--first-seed", type=int, default=100)
    a = ap.parse_args()
    from oracle import refhook, sroracle
    from pindel_b200 import bam, bgzf
    from tests import sr_fuzz_helpers as F
    from tests.emul import bam_emul, sr_emul
    from tests.test_sr_emulator import rows_of_product
    assert refhook.available(), "needs oracle/_ref (build in the container that has /root/reference)"
    lib, bl = sr_emul.build(), bam_emul.build()
    tmp = tempfile.mkdtemp()
    rng = np.random.default_rng(2)
    fa = os.path.join(tmp, "g.fa")
    refhook.write_fasta(fa, [(n, rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), l)) for n, l in F.REFS])
    h = refhook.RefHook(fa, [], workdir=tmp)
    path = os.path.join(tmp, "x.bam")
    total = 0
    for seed in range(a.first_seed, a.first_seed + a.rounds):
        stream = F.fuzz_stream(seed, 800)
        open(path, "wb").write(bgzf.compress(stream))
        open(path + ".bai", "wb").write(b"BAI\x01\x00\x00\x00\x00")
        table = bam.serial_index(stream)
        recs, _ = bam_emul.index(bl, stream)
        for tid, beg, end, isz, min_q in F.QUERIES:
            cand, refs = h.sr_candidates(path, tid, beg, end, isz, min_q)
            oc, orf = sroracle.candidates(stream, table, tid, beg, end, isz, F.REFS[tid][1], min_anchor_quality=min_q)
            assert np.array_equal(cand, oc) and np.array_equal(refs, orf), ("oracle vs reference", seed, tid, beg, end)
            cands, seqs, rr, _, err = sr_emul.candidates(lib, stream, recs, sr_emul.query(tid, beg, end, isz, F.REFS[tid][1], min_q), seed, 0, 0xFFF << 20)
            c, r = rows_of_product(stream, recs, cands, seqs, rr, isz)
            assert err == 0 and np.array_equal(c, oc) and np.array_equal(r, orf), ("device functions vs reference", seed, tid, beg, end)
            total += oc.shape[0]
    print("ok: %d rounds, %d candidates compared three ways" % (a.rounds, total))


if __name__ == "__main__":
    main()
