"""scripts/cpu_fuzz_campaign.py -- a longer randomized differential run of the device programs' host builds (tests/emul/)
against the REFERENCE ITSELF (oracle/_ref/libpindel_ref.so) and zlib, for the stages whose reference functions need no
initialised genome: read-pair clustering (both variants), per-box ordering, BGZF inflate.  Not part of the pytest suite (minutes of CPU);
run it after touching pg_rp.cuh / pg_rp_host.h / pg_boxsort.cuh / pg_bgzf.cuh:
    python scripts/cpu_fuzz_campaign.py [--rounds N]"""
from __future__ import annotations

import argparse
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=30)
    ap.add_argument("--seed", type=int, default=2026)
    a = ap.parse_args()
    from oracle import boxsortoracle as B, refhook, rporacle
    from pindel_b200 import bgzf, synth
    from tests import boxsort_helpers as H
    from tests.emul import bgzf_emul, boxsort_emul, rp_emul
    from tests.golden.make_rp_golden import rows_from_table
    assert refhook.available(), "needs oracle/_ref (build in the container that has /root/reference)"
    rp_lib, box_lib, z_lib = rp_emul.build(), boxsort_emul.build(), bgzf_emul.build()
    rng = np.random.default_rng(a.seed)
    for r in range(a.rounds):
        # read-pair clustering: spans from dense to sparse, 1-3 libraries, with and without duplicates
        libs = [((100, 250),), ((150, 500),), ((100, 250), (150, 500)), ((76, 200), (100, 350), (150, 600))][int(rng.integers(0, 4))]
        rows = synth.make_discordant_pairs(int(rng.integers(1, 200)), 5000 + r + 1000 * (a.seed - 2026), span=int(rng.choice([60_000, 400_000, 5_000_000, 60_000_000])),
                                           background=int(rng.integers(0, 300)), max_support=int(rng.integers(2, 80)), libraries=libs,
                                           both_mates=float(rng.choice([0.0, 0.5, 0.9])))
        table, tags, _ = refhook.rp_cluster(rows)
        srt = rows_from_table(table)
        _, got = rp_emul.cluster(rp_lib, srt)
        for col, key in ((7, "pos_a"), (8, "pos_a1"), (9, "pos_b"), (10, "pos_b1"), (11, "visited"), (12, "report"), (13, "n_identical")):
            assert np.array_equal(got[key].astype(np.int64), table[:, col]), ("rp", r, key)
        holder = {}
        for i, tg in enumerate(tags):
            for x in tg:
                holder[x] = i
        assert all(holder[int(table[i, 0])] == int(got["merged_into"][i]) for i in range(srt.size)), ("rp tags", r)
        # the inter-chromosomal variants
        irows, ica, icb = synth.make_interchr_pairs(int(rng.integers(1, 120)), 6000 + r + 1000 * (a.seed - 2026), n_chr=int(rng.integers(2, 7)),
                                                    max_support=int(rng.integers(2, 70)), background=int(rng.integers(0, 200)))
        itable, itags, isa, isb = refhook.rp_cluster_interchr(irows, ica, icb)
        _, igot = rp_emul.cluster_interchr(rp_lib, rows_from_table(itable), isa, isb)
        for col, key in ((7, "pos_a"), (8, "pos_a1"), (9, "pos_b"), (10, "pos_b1"), (11, "visited"), (12, "report"), (13, "n_identical")):
            assert np.array_equal(igot[key].astype(np.int64), itable[:, col]), ("rp interchr", r, key)
        # per-box ordering
        reads, box_off, idx = B.make_boxes(int(rng.integers(1, 120)), 7000 + r + 1000 * (a.seed - 2026), max_reads=int(rng.integers(3, 120)), heavy=int(rng.integers(0, 2)))
        r_idx, r_unique = B.reference_sort_boxes(reads, box_off, idx)
        for warp in (False, True):
            order, non = boxsort_emul.sort_boxes(box_lib, reads[idx], box_off, warp)
            p_idx, p_unique = H.apply(reads, box_off, idx, order, non)
            assert np.array_equal(p_idx, r_idx) and np.array_equal(p_unique, r_unique), ("boxsort", r, warp)
        # BGZF inflate
        kind = int(rng.integers(0, 4))
        n = int(rng.integers(0, 400_000))
        payload = (synth.make_bam_like_bytes(n, 9000 + r + 1000 * (a.seed - 2026)) if kind == 0 else bytes(rng.integers(0, int(rng.integers(2, 257)), n, dtype=np.uint8)) if kind == 1
                   else bytes(rng.choice(40, size=n, p=(lambda p: p / p.sum())(0.6 ** np.arange(40))).astype(np.uint8)) if kind == 2
                   else (b"ACGT" * 50 + bytes(rng.integers(0, 256, 30, dtype=np.uint8))) * (n // 230 + 1))
        comp = bgzf.compress(payload, level=int(rng.integers(0, 10)), block=int(rng.choice([0x4000, 0x7000])) if kind == 1 else 0xff00,
                             strategy=int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED])))
        got, status = bgzf_emul.inflate(z_lib, comp)
        assert not status.any() and got == payload, ("bgzf", r, kind)
        print("round %d ok: %d + %d pairs, %d boxes, %d payload bytes" % (r, rows.size, irows.size, box_off.size - 1, len(payload)), flush=True)
    print("CAMPAIGN-OK")


if __name__ == "__main__":
    main()
