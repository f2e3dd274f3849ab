"""scripts/bam_index_bench.py -- measurement of BGZF inflate + BAM record index on one GPU,
with the serial walk (the same functions compiled for the host, one segment = one chain walk in C) beside it.

    python scripts/bam_index_bench.py [--mb N]

Prints one JSON line."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    from pindel_b200 import api, bam, bgzf
    from tests.emul import bam_emul
    # one chunk of distinct records, tiled (the index does not care that positions repeat); chunk = whole BGZF blocks
    one = bam.synthetic_stream(60_000, 1, refs=(("chr1", 2_000_000_000),), discordant=0.05, switch_ref=False, unplaced=0.0)
    first, _ = bam.first_record_offset(one)
    hdr, chunk = one[:first], one[first:]
    tiles = max(1, (a.mb << 20) // len(chunk))
    comp = bgzf.compress(hdr, eof=False) + bgzf.compress(chunk, eof=False) * tiles + bgzf.EOF_MARKER
    stream_len = len(hdr) + len(chunk) * tiles
    n_expected = bam.serial_index(one).size * tiles
    buf = np.frombuffer(comp, dtype=np.uint8)
    pg = api.PatternGrowth()
    assert pg.bgzf_inflate(buf, to_host=False) == stream_len
    recs, _, _ = pg.bam_index()
    assert recs.size == n_expected
    one_idx = bam.serial_index(one)
    assert np.array_equal(recs["off"][:one_idx.size], one_idx["off"]) and np.array_equal(recs["end"][-one_idx.size:], one_idx["end"])
    out = np.zeros(recs.size, dtype=bam.REC_DT)
    t_inf, t_idx, k_idx = [], [], []
    for _ in range(a.reps):
        t0 = time.perf_counter(); pg.bgzf_inflate(buf, to_host=False); t_inf.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); pg.bam_index(recs=out); t_idx.append(time.perf_counter() - t0)
        k_idx.append(pg.bam_last_ms)
    res = {"stage": "BGZF inflate + BAM record index", "inflated_bytes": int(stream_len), "compressed_bytes": int(buf.size), "records": int(recs.size),
           "segments": pg.bam_last_segments()[0], "segments_walked_again": pg.bam_last_segments()[1],
           "inflate_e2e_ms": 1e3 * float(np.median(t_inf)), "index_e2e_ms": 1e3 * float(np.median(t_idx)), "index_device_ms": float(np.median(k_idx)),
           "index_records_per_s_e2e": recs.size / float(np.median(t_idx)), "index_GBps_of_stream": stream_len / float(np.median(t_idx)) / 1e9,
           "d2h_bytes": int(out.nbytes)}
    # discordant-pair discovery over the whole first reference (the tiles repeat positions, so every tile contributes)
    rows, src, irows, isrc, scanned = pg.bam_rp_discovery(0, 0, 2_000_000_000, 500)
    t_d, k_d = [], []
    for _ in range(a.reps):
        t0 = time.perf_counter(); pg.bam_rp_discovery(0, 0, 2_000_000_000, 500); t_d.append(time.perf_counter() - t0)
        k_d.append(pg.bam_last_ms)
    res["discovery"] = {"records_scanned": int(scanned), "rows": int(rows.size), "inter_rows": int(irows.size), "e2e_ms": 1e3 * float(np.median(t_d)),
                        "device_ms": float(np.median(k_d)), "records_per_s_e2e": scanned / float(np.median(t_d)),
                        "note": "two passes through the C ABI (sizes, then rows), as the wrapper calls it"}
    # the serial walk in C over the same stream (host memory)
    lib = bam_emul.build()
    stream = np.frombuffer(hdr + chunk * tiles, dtype=np.uint8)
    stats = np.zeros(3, dtype=np.uint64)
    lib.bamemu_index(stream.ctypes.data, stream.size, stream.size, out.ctypes.data, out.size, stats.ctypes.data)
    t0 = time.perf_counter()
    n = lib.bamemu_index(stream.ctypes.data, stream.size, stream.size, out.ctypes.data, out.size, stats.ctypes.data)
    dt = time.perf_counter() - t0
    assert n == recs.size
    res["cpu_serial_walk"] = {"ms": 1e3 * dt, "records_per_s": n / dt, "threads": 1,
                              "note": "count pass + fill pass of the same walk in C (tests/emul/bam_emul.cpp, one segment)"}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
