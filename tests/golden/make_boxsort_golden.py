"""Golden vectors of row N4 (per-box read ordering + duplicate marking) made by the REFERENCE ITSELF: its own bubblesortReads
and markDuplicates (src/reporter.cpp:932-972) through oracle/ref_hook.cpp on seeded synthetic boxes.
    python -m tests.golden.make_boxsort_golden"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CASES = {"small": dict(n_boxes=300, seed=1), "ties": dict(n_boxes=200, seed=2, max_reads=80), "heavy": dict(n_boxes=40, seed=3, heavy=6)}


def main():
    from oracle import boxsortoracle as B
    for name, kw in CASES.items():
        reads, box_off, idx = B.make_boxes(**kw)
        out_idx, unique = B.reference_sort_boxes(reads, box_off, idx)
        np.savez_compressed(os.path.join(HERE, "boxsort_%s.npz" % name), sorted_idx=out_idx, unique=unique)
        print(name, reads.size, "reads in", box_off.size - 1, "boxes,", int((reads["unique"] == 1).sum() - unique.sum()), "marked non-unique,",
              int((out_idx != idx).sum()), "positions moved")


if __name__ == "__main__":
    main()
