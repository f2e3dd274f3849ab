"""Golden vectors of row N2 (discordant-pair clustering) made by the REFERENCE ITSELF: its own sort + ModifyRP + Summarize
(oracle/ref_hook.cpp pgref_rp_cluster) on seeded synthetic pair sets.  Run in the container where /root/reference exists:
    python -m tests.golden.make_rp_golden
The .npz holds the reference's output table (which contains the inputs in the order ModifyRP's sort leaves them) and the
Tags of every row."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name -> make_discordant_pairs arguments
CASES = {
    "sparse": dict(n_events=150, seed=101, background=400),
    "dense": dict(n_events=400, seed=102, span=1_000_000, background=100, max_support=60),     # chains of overlapping intervals
    "tiny": dict(n_events=1, seed=103, max_support=6),
    "one_library": dict(n_events=200, seed=104, libraries=((100, 250),), span=5_000_000),
    "no_duplicates": dict(n_events=200, seed=105, both_mates=0.0, span=5_000_000, background=50),
}


# inter-chromosomal variants: name -> make_interchr_pairs arguments
INTER_CASES = {
    "inter_small": dict(n_events=60, seed=201, background=40),
    "inter_two_chr": dict(n_events=120, seed=202, n_chr=2, max_support=50),
    "inter_many_chr": dict(n_events=150, seed=203, n_chr=6, background=200),
}


def rows_from_table(table):
    from oracle import rporacle
    rows = np.zeros(table.shape[0], dtype=rporacle.RP_DT)
    rows["pos_a"], rows["pos_b"], rows["da"], rows["db"] = table[:, 1], table[:, 2], table[:, 3], table[:, 4]
    rows["read_length"], rows["insert_size"] = table[:, 5], table[:, 6]
    return rows


def main():
    from oracle import refhook
    from pindel_b200 import synth
    for name, kw in CASES.items():
        rows = synth.make_discordant_pairs(**kw)
        table, tags, secs = refhook.rp_cluster(rows)
        np.savez_compressed(os.path.join(HERE, "rp_%s.npz" % name), table=table.astype(np.int64),
                            tag_counts=np.array([len(t) for t in tags], np.int32),
                            tag_values=np.array([x for t in tags for x in t], np.int32))
        print(name, len(rows), "rows,", int(table[:, 12].sum()), "reported,", int(table[:, 11].sum()), "visited, %.3f s" % sum(secs))
    for name, kw in INTER_CASES.items():
        rows, ca, cb = synth.make_interchr_pairs(**kw)
        table, tags, sa, sb = refhook.rp_cluster_interchr(rows, ca, cb)
        np.savez_compressed(os.path.join(HERE, "rp_%s.npz" % name), table=table.astype(np.int64), chr_a=sa, chr_b=sb,
                            tag_counts=np.array([len(t) for t in tags], np.int32), tag_values=np.array([x for t in tags for x in t], np.int32))
        print(name, len(rows), "rows,", int(table[:, 12].sum()), "reported,", int(table[:, 11].sum()), "visited")


if __name__ == "__main__":
    main()
