"""Shared by the CPU and GPU tests of row N4: turn the product's per-box (order, nonunique) into the reference's view
(reordered read indices, UniqueRead per read)."""
import numpy as np


def apply(reads, box_off, idx, order, nonunique):
    out_idx = idx.copy()
    unique = reads["unique"].copy()
    for b in range(len(box_off) - 1):
        a, e = int(box_off[b]), int(box_off[b + 1])
        out_idx[a:e] = idx[a:e][order[a:e]]
        unique[out_idx[a:e][nonunique[a:e] != 0]] = 0
    return out_idx, unique
