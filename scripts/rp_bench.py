"""scripts/rp_bench.py -- measurement of the read-pair clustering stage on one GPU with the
reference's own ModifyRP + Summarize timed beside it (one thread: its OpenMP loop is a data race).

    python scripts/rp_bench.py [--events N] [--span BASES] [--support K]

Prints one JSON line: rows, pairs compared by the reference's double loop (rows^2 per strand group), device and end-to-end
times of pg_rp_modify / pg_rp_summarize, reference seconds, identical-output check."""
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
    ap.add_argument("--events", type=int, default=2000)
    ap.add_argument("--span", type=int, default=60_000_000)
    ap.add_argument("--support", type=int, default=40)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    from pindel_b200 import api, rp, synth
    rows = synth.make_discordant_pairs(a.events, 777, span=a.span, background=a.events, max_support=a.support)
    res = {"stage": "read-pair clustering (ModifyRP + Summarize)", "rows": int(rows.size)}
    table = None
    from oracle import refhook
    if refhook.available() and not a.no_cpu:
        table, _, secs = refhook.rp_cluster(rows)
        from tests.golden.make_rp_golden import rows_from_table
        rows = rows_from_table(table)            # the order the reference's own sorts left
        res["cpu_reference"] = {"modify_s": secs[0], "summarize_s": secs[1], "threads": 1,
                                "note": "the reference's own sort + ModifyRP + Summarize (oracle/_ref/libpindel_ref.so)"}
    else:
        rows = rows[rp.compare2rp_order(rows)]
    keys = (rows["da"].astype(np.uint32) << 8) | rows["db"]
    res["groups"] = {chr(int(k) >> 8) + chr(int(k) & 255): int(c) for k, c in zip(*np.unique(keys, return_counts=True))}
    res["pairs_compared_by_reference"] = int(rows.size) ** 2
    pg = api.PatternGrowth()
    mod = pg.rp_modify(rows)
    fin = pg.rp_summarize(rows, mod)
    tm, ts, km, ks = [], [], [], []
    for _ in range(a.reps):
        t0 = time.perf_counter(); mod = pg.rp_modify(rows); tm.append(time.perf_counter() - t0); km.append(pg.rp_last_kernel_ms)
        t0 = time.perf_counter(); fin = pg.rp_summarize(rows, mod); ts.append(time.perf_counter() - t0); ks.append(pg.rp_last_kernel_ms)
    res.update({"modify_e2e_ms": 1e3 * float(np.median(tm)), "modify_kernel_ms": float(np.median(km)),
                "summarize_e2e_ms": 1e3 * float(np.median(ts)), "summarize_kernel_ms": float(np.median(ks)),
                "rows_per_s_e2e": rows.size / (float(np.median(tm)) + float(np.median(ts))),
                "h2d_bytes": int(rows.nbytes * 2 + mod.nbytes + rows.size * 4), "d2h_bytes": int(mod.nbytes * 2),
                "reported_events": int(fin["report"].sum())})
    pg.rp_modify(rows)
    res["classes"] = dict(zip(("one_warp_each", "one_cta_each", "rows_in_classes"), pg.rp_last_classes()))
    if table is not None:
        same = all(np.array_equal(fin[k].astype(np.int64), table[:, c]) for c, k in
                   ((7, "pos_a"), (8, "pos_a1"), (9, "pos_b"), (10, "pos_b1"), (11, "visited"), (12, "report"), (13, "n_identical")))
        res["identical_to_reference"] = bool(same)
        res["speedup_e2e_vs_cpu"] = (res["cpu_reference"]["modify_s"] + res["cpu_reference"]["summarize_s"]) / \
                                    (float(np.median(tm)) + float(np.median(ts)))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
