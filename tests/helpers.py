"""Shared helpers of the test-suite: run the oracle, compare result sets."""
from __future__ import annotations

import numpy as np

from oracle import pgoracle, refhook
from pindel_b200 import runs as R

KEYS = ("close_raw", "close", "far")


def oracle_results(chr_seqs, batch, chr_id=None, regions=None, region_off=None, threads=8, **params):
    """Run the C restatement (oracle/pg_oracle.c) -> list of per-read dicts."""
    P = pgoracle.default_params(**params)
    blob, off = batch.blob()
    stream, _ = pgoracle.search(P, chr_seqs, blob, off, batch.matched_d, batch.matched_rel_pos, batch.insert_size,
                                chr_id=chr_id, regions=regions, region_off=region_off, threads=threads)
    return refhook.parse_stream(stream, batch.n)


def compare(expected, got, what=""):
    """expected/got: lists of per-read dicts (flip, L, close_raw, close, far).  Raises AssertionError with the first diffs."""
    assert len(expected) == len(got)
    bad = []
    for i, (a, b) in enumerate(zip(expected, got)):
        ok = a["flip"] == b["flip"] and a["L"] == b["L"] and all(np.array_equal(a[k], b[k]) for k in KEYS)
        if not ok:
            bad.append(i)
    if bad:
        i = bad[0]
        a, b = expected[i], got[i]
        msg = ["%s: %d of %d reads differ; first = read %d" % (what, len(bad), len(expected), i),
               " flip exp/got %d/%d  L %d/%d" % (a["flip"], b["flip"], a["L"], b["L"])]
        for k in KEYS:
            if not np.array_equal(a[k], b[k]):
                msg.append(" %s exp %s" % (k, a[k][:5].tolist()))
                msg.append(" %s got %s" % (k, b[k][:5].tolist()))
        raise AssertionError("\n".join(msg))


def batch_kwargs(batch, chr_id=None):
    blob, off = batch.blob()
    d = dict(seq_blob=blob, seq_off=off, matched_d=batch.matched_d, matched_rel_pos=batch.matched_rel_pos,
             insert_size=batch.insert_size)
    if chr_id is not None:
        d["chr_id"] = chr_id
    return d
