"""scripts/bd_prepare_bench.py -- CPU-only: what BDData::loadRegion costs in the reference (its position sweep over a 10 Mbp window,
through oracle/_ref/libpindel_ref.so) beside the preparation the library does instead (pgbd_prepare, host build of pg_bd_host.h:
two sorted arrays), on the same 4000 events; and that both give the same clusters for 100 000 close-end positions.
Prints one JSON line.  Needs oracle/_ref (the container that has /root/reference)."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import bdoracle, refhook
    from tests import bd_helpers as B
    from tests.emul import bd_emul
    rng = np.random.default_rng(5)
    tmp = tempfile.mkdtemp()
    fa = os.path.join(tmp, "g.fa")
    refhook.write_fasta(fa, [(n, rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), l)) for n, l in zip(B.NAMES, B.LENS)])
    hook = refhook.RefHook(fa, [], workdir=tmp)
    W = (100_000, 100_000 + 10_000_000)
    n_ev = 4000
    a = rng.integers(W[0], W[1], n_ev)
    b = rng.integers(W[0], W[1], n_ev)
    ev = np.stack([np.zeros(n_ev, int), a, a + rng.integers(0, 300, n_ev), np.zeros(n_ev, int), b, b + rng.integers(0, 300, n_ev)], 1)
    q = rng.integers(W[0], W[1], 100_000).astype(np.uint32)
    t_ref = []
    for _ in range(5):
        t = time.perf_counter()
        bdoracle.reference_clusters(hook, ev, 0, W[0], W[1], q[:1])       # event construction + sort + loadRegion + one lookup
        t_ref.append(time.perf_counter() - t)
    lib = bd_emul.build()
    ev_sorted = np.array(bdoracle.sort_on_first(bdoracle.normalise(ev), B.NAMES), dtype=np.int64).reshape(-1, 6)
    t_prep = []
    for _ in range(5):
        t = time.perf_counter()
        bd_emul.clusters(lib, ev_sorted, B.NAMES, 0, W[0], W[1], q[:1])       # pgbd_prepare + one lookup
        t_prep.append(time.perf_counter() - t)
    exp = bdoracle.reference_clusters(hook, ev, 0, W[0], W[1], q)
    got = bd_emul.clusters(lib, ev_sorted, B.NAMES, 0, W[0], W[1], q)
    print(json.dumps({"window_positions": W[1] - W[0] + 6000, "events": n_ev, "reference_loadRegion_ms": round(1e3 * float(np.median(t_ref)), 2),
                      "reference_mask_bytes": 4 * (W[1] - W[0] + 6001), "library_prepare_ms": round(1e3 * float(np.median(t_prep)), 3),
                      "library_array_bytes": int(n_ev * (4 + 4 + 24)), "lookups_compared": int(q.size),
                      "non_empty_clusters": int(sum(1 for c in exp if c)), "equal": bool(exp == got)}))


if __name__ == "__main__":
    main()
