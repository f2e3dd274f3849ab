"""Kernel time of pg_box_sort_kind for every reporter's rules on 20 000 boxes (three of them with hundreds of reads), beside
oracle/boxsort_oracle.c's restatement of the reference's loops on one host thread (the reporters run them serially).
Prints one JSON line per kind."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import boxsortoracle as B          # noqa: E402  (the checker and the CPU leg of this measurement)
from pindel_b200 import api                    # noqa: E402

NAMES = {0: "D/TD", 1: "SI", 2: "DI", 3: "INV", 4: "INV_NT"}


def main():
    pg = api.PatternGrowth()
    for kind in NAMES:
        if kind <= 1:
            reads, box_off, idx = B.make_boxes(20_000, 70 + kind, max_reads=30, heavy=3)
        else:
            reads, box_off, idx = B.make_boxes_kind(kind, 20_000, 70 + kind, max_reads=30, heavy=3)
        keys = reads[idx]
        pg.box_sort_kind(kind, keys, box_off)                  # warm-up
        ms, e2e = [], []
        for _ in range(5):
            t = time.perf_counter()
            order, non, ev = pg.box_sort_kind(kind, keys, box_off)
            e2e.append((time.perf_counter() - t) * 1e3)
            ms.append(pg.box_sort_last_kernel_ms)
        t = time.perf_counter()
        o_idx, o_unique, o_event = B.sort_boxes_kind(kind, reads, box_off, idx)
        cpu_ms = (time.perf_counter() - t) * 1e3
        same = bool(np.array_equal(idx[np.concatenate([order[box_off[b]:box_off[b + 1]] + box_off[b] for b in range(len(box_off) - 1)])], o_idx)
                    and np.array_equal(ev, o_event))
        print(json.dumps({"kind": NAMES[kind], "boxes": int(len(box_off) - 1), "reads": int(idx.size), "kernel_ms": round(float(np.median(ms)), 3),
                          "call_ms_host_arrays": round(float(np.median(e2e)), 3), "oracle_loops_one_thread_ms": round(cpu_ms, 1),
                          "equal_to_oracle": same}), flush=True)
    pg.close()


if __name__ == "__main__":
    main()
