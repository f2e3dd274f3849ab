"""scripts/boxsort_bench.py -- measurement of the per-box read ordering + duplicate marking on
one GPU with the reference's own bubblesortReads + markDuplicates (one thread, as the reporters call them) beside it.

    python scripts/boxsort_bench.py [--boxes N] [--max-reads K]

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
    ap.add_argument("--boxes", type=int, default=20_000)
    ap.add_argument("--max-reads", type=int, default=40)
    ap.add_argument("--heavy", type=int, default=4)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    from oracle import boxsortoracle as B, refhook
    from pindel_b200 import api
    from tests import boxsort_helpers as H
    reads, box_off, idx = B.make_boxes(a.boxes, 11, max_reads=a.max_reads, heavy=a.heavy)
    keys = reads[idx]
    pg = api.PatternGrowth()
    order, non = pg.box_sort(keys, box_off)
    wall, kern = [], []
    for _ in range(a.reps):
        t0 = time.perf_counter(); order, non = pg.box_sort(keys, box_off); wall.append(time.perf_counter() - t0)
        kern.append(pg.box_sort_last_kernel_ms)
    sizes = np.diff(box_off.astype(np.int64))
    res = {"stage": "per-box read ordering + duplicate marking", "boxes": int(a.boxes), "reads": int(keys.size), "largest_box": int(sizes.max()),
           "pair_comparisons": int((sizes * (sizes - 1)).sum()), "e2e_ms": 1e3 * float(np.median(wall)), "kernel_ms": float(np.median(kern)),
           "reads_per_s_e2e": keys.size / float(np.median(wall)), "h2d_bytes": int(keys.nbytes + box_off.nbytes), "d2h_bytes": int(order.nbytes + non.nbytes)}
    p_idx, p_unique = H.apply(reads, box_off, idx, order, non)
    if refhook.available():
        t0 = time.perf_counter()
        r_idx, r_unique = B.reference_sort_boxes(reads, box_off, idx)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        B.reference_sort_boxes(reads[:1], np.array([0, 1], dtype=np.uint32), np.array([0], dtype=np.uint32))
        res["cpu_reference"] = {"seconds": dt, "threads": 1, "note": "the reference's bubblesortReads + markDuplicates per box (includes building the SPLIT_READs: "
                                "%.3f s for one read)" % (time.perf_counter() - t0)}
        res["identical_to_reference"] = bool(np.array_equal(p_idx, r_idx) and np.array_equal(p_unique, r_unique))
        res["speedup_e2e_vs_cpu"] = dt / float(np.median(wall))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
