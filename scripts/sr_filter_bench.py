"""scripts/sr_filter_bench.py -- measurement of the reader's split-read filter on one GPU (rest of row
N3): file bytes -> pg_bgzf_inflate -> pg_bam_index -> pg_bam_sr_candidates, with the reference's own ReadInBamReads_SR (through
oracle/_ref/libpindel_ref.so: zlib inflate, record walk, per-record loop, one thread) timed beside it on a bounded prefix.

    python scripts/sr_filter_bench.py [--pairs N] [--ref-pairs M]

Prints one JSON line.  (Written before the stage's first hardware run)"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

REFS = (("chr1", 400_000), ("chr2", 300_000))


def tiled_pairs(n_pairs: int, chunk_pairs: int = 20_000, seed: int = 5):
    """A stream of about n_pairs named pairs: one seeded, coordinate-sorted chunk of chr1 records, repeated with the tile number
    written into every read name (names stay unique, so pairs never join across tiles).  Positions start over in every tile;
    the benchmark's window is the whole reference, so the iterator's stop rule never fires.  -> (stream, records)"""
    from pindel_b200 import bam
    from tests.golden import make_sr_golden as G
    one = G.paired_stream(chunk_pairs, seed, refs=REFS[:1])
    first, _ = bam.first_record_offset(one)
    table = bam.serial_index(one)
    body = np.frombuffer(one[first:], dtype=np.uint8)
    name_at = (table["off"].astype(np.int64) - first + 36)            # "pair_<i>": overwrite "pair" with the tile number
    tiles = max(1, n_pairs // chunk_pairs)
    out = [one[:first]]
    for t in range(tiles):
        b = body.copy()
        digits = np.frombuffer(b"%04d" % t, dtype=np.uint8)
        for k in range(4):
            b[name_at + k] = digits[k]
        out.append(b.tobytes())
    return b"".join(out), int(table.size) * tiles


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1_000_000)
    ap.add_argument("--ref-pairs", type=int, default=100_000, help="size of the prefix the reference's reader is timed on")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    from pindel_b200 import api, bam, bgzf
    stream, n_records = tiled_pairs(a.pairs)
    comp = np.frombuffer(bgzf.compress(stream), dtype=np.uint8)
    pg = api.PatternGrowth()
    q = (0, 0, REFS[0][1], 500, REFS[0][1])
    assert pg.bgzf_inflate(comp, to_host=False) == len(stream)
    recs, _, _ = pg.bam_index()
    assert recs.size == n_records
    cands, seqs, refs, scanned = pg.bam_sr_candidates(*q)
    t_all, t_f, k_f = [], [], []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        pg.bgzf_inflate(comp, to_host=False)
        pg.bam_index(recs=recs)
        t1 = time.perf_counter()
        pg.bam_sr_candidates(*q)
        t2 = time.perf_counter()
        t_all.append(t2 - t0); t_f.append(t2 - t1); k_f.append(pg.bam_last_ms)
    res = {"stage": "split-read filter (ReadInBamReads_SR)", "records": int(n_records), "records_delivered": int(scanned), "candidates": int(cands.size),
           "ref_supporting_reads": int(refs.size), "stride": int(seqs.shape[1]), "inflated_bytes": len(stream), "compressed_bytes": int(comp.size),
           "filter_e2e_ms": 1e3 * float(np.median(t_f)), "filter_device_ms": float(np.median(k_f)),
           "file_to_candidates_e2e_ms": 1e3 * float(np.median(t_all)), "records_per_s_filter_e2e": scanned / float(np.median(t_f)),
           "records_per_s_file_to_candidates": scanned / float(np.median(t_all)),
           "d2h_bytes": int(cands.nbytes + seqs.nbytes + refs.nbytes),
           "note": "filter = two passes through the C ABI (sizes, then rows); device_ms = kernels of the second pass"}
    pg.close()
    # the reference's own reader on a prefix of the same stream, and the equality of what both produce there
    from oracle import refhook
    if refhook.available():
        from tests.test_sr_emulator import rows_of_product
        small, n_small = tiled_pairs(a.ref_pairs)
        tmp = tempfile.mkdtemp()
        rng = np.random.default_rng(1)
        fa = os.path.join(tmp, "g.fa")
        refhook.write_fasta(fa, [(n, rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), l)) for n, l in REFS])
        path = os.path.join(tmp, "x.bam")
        open(path, "wb").write(bgzf.compress(small))
        open(path + ".bai", "wb").write(b"BAI\x01\x00\x00\x00\x00")
        h = refhook.RefHook(fa, [], workdir=tmp)
        t0 = time.perf_counter()
        rc, rr = h.sr_candidates(path, 0, 0, REFS[0][1], 500, 0)
        dt = time.perf_counter() - t0
        pg = api.PatternGrowth()
        pg.bgzf_inflate(np.frombuffer(bgzf.compress(small), dtype=np.uint8), to_host=False)
        srecs, _, _ = pg.bam_index()
        c2, s2, r2, sc2 = pg.bam_sr_candidates(*q)
        c, r = rows_of_product(small, srecs, c2, s2, r2, 500)
        res["cpu_reference"] = {"records": int(n_small), "ms": 1e3 * dt, "records_per_s": n_small / dt, "threads": 1, "equal_to_device": bool(
            np.array_equal(c, rc) and np.array_equal(r, rr)), "note": "ReadInBamReads_SR through the hook: open, zlib inflate, record walk, per-record loop"}
        pg.close()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
