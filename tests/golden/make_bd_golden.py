"""tests/golden/make_bd_golden.py -- regenerates bd_clusters.npz FROM THE REFERENCE ITSELF.

For seeded event lists (tests/bd_helpers.py) it runs the reference's own BDData::loadRegion and
getCorrespondingSearchWindowCluster (src/bddata.cpp:852-967) through oracle/_ref/libpindel_ref.so (oracle/ref_hook_bd.cpp,
pgref_bd_clusters) and stores events, window, query positions and the clusters the reference returned.

    python tests/golden/make_bd_golden.py          # needs /root/reference and `make -C oracle ref`
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

SEEDS = list(range(1, 13))


def main():
    from oracle import bdoracle, refhook
    from tests import bd_helpers as B
    rng = np.random.default_rng(5)
    tmp = tempfile.mkdtemp()
    fa = os.path.join(tmp, "g.fa")
    refhook.write_fasta(fa, [(n, rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), l)) for n, l in zip(B.NAMES, B.LENS)])
    hook = refhook.RefHook(fa, [], workdir=tmp)
    out = {"names": np.array(B.NAMES), "seeds": np.array(SEEDS)}
    for seed in SEEDS:
        chr_id, window, ev, q = B.case(seed, seed % 3 == 0, n_queries=300)
        exp = bdoracle.reference_clusters(hook, ev, chr_id, window[0], window[1], q)
        cnt = np.array([len(c) for c in exp], dtype=np.int32)
        flat = np.array([w for c in exp for w in c], dtype=np.int64).reshape(-1, 3)
        out["ev_%d" % seed] = ev.astype(np.int32)
        out["win_%d" % seed] = np.array([chr_id, window[0], window[1]], dtype=np.int64)
        out["q_%d" % seed] = q
        out["cnt_%d" % seed] = cnt
        out["reg_%d" % seed] = flat.astype(np.uint32)
        print("seed %d: %d events, %d queries, %d windows in %d non-empty clusters" % (seed, ev.shape[0], q.size, flat.shape[0], int((cnt > 0).sum())))
    np.savez_compressed(os.path.join(HERE, "bd_clusters.npz"), **out)


if __name__ == "__main__":
    main()
