"""Shared inputs of the BreakDancer-mask tests: event lists that reach the corners of the reference's sweep."""
import numpy as np

SPACER = 100000
NAMES = ["chr10", "chr2", "chrA"]          # string order differs from index order on purpose ("chr10" < "chr2")
LENS = [400_000, 300_000, 200_000]


def random_events(rng, n, chr_id, window, dense=False):
    """n events whose first coordinate falls around `window` on chr_id (two positions per coordinate as UpdateBD builds them,
    src/bddata.cpp:706-707): starts left of the window, clusters of nearly equal positions, long second spans, partners on
    other chromosomes, exact duplicates."""
    ws, we = window
    ev = []
    for _ in range(n):
        if dense:
            a = int(rng.integers(ws + 2000, ws + 6000))
        else:
            a = int(rng.integers(max(0, ws - 3600), we + 3600))
        a2 = a + int(rng.choice([0, 0, 1, 50, 300, 2000]))
        cb = int(rng.choice([chr_id, chr_id, chr_id, (chr_id + 1) % 3, (chr_id + 2) % 3]))
        b = int(rng.integers(SPACER, SPACER + LENS[cb] - 3000))
        b2 = b + int(rng.choice([0, 0, 1, 100, 350, 1500]))
        ev.append((chr_id, a, a2, cb, b, b2))
        if rng.random() < 0.15:
            ev.append(ev[-1])                                   # exact duplicate
        if rng.random() < 0.3:
            ev.append((cb, b, b2, chr_id, a, a2))               # the mirrored event UpdateBD also adds
        if rng.random() < 0.2:                                  # a neighbour within the window span: overlapping windows
            ev.append((chr_id, a + int(rng.integers(0, 250)), a2 + int(rng.integers(0, 400)), cb, b + int(rng.integers(0, 500)), b2 + int(rng.integers(0, 500))))
    return np.array(ev, dtype=np.int64).reshape(-1, 6)


def queries(rng, window, n):
    ws, we = window
    q = rng.integers(max(0, ws - 3500), we + 3500, n)
    edge = np.array([max(0, ws - 3000), max(0, ws - 3000) + 1, ws, we, we + 2999, we + 3002, max(0, ws - 3001)])
    q = np.concatenate([q, edge])
    # the reference's mask has one entry more than its sweep writes and the range check lets one more position through
    # (src/bddata.cpp:896, :904, :952): close ends at window end + 3000 and + 3001 read garbage there -- not asked
    q = q[(q != we + 3000) & (q != we + 3001)]
    return q.astype(np.uint32)


def case(seed, dense, n_queries=600):
    """-> (chromosome, window, events [n, 6], query positions) of one seeded case"""
    rng = np.random.default_rng(500 + seed)
    chr_id = int(rng.integers(0, 3))
    ws = int(rng.choice([SPACER, SPACER + 500, SPACER + 40_000, 1000, 0]))
    window = (ws, ws + int(rng.choice([5_000, 20_000, 60_000])))
    ev = random_events(rng, int(rng.choice([0, 1, 3, 40, 200])) if not dense else 200, chr_id, window, dense)
    other = random_events(rng, 20, (chr_id + 1) % 3, window)
    ev = np.concatenate([ev, other])
    return chr_id, window, ev, queries(rng, window, n_queries)


def load_golden():
    """tests/golden/bd_clusters.npz (written by tests/golden/make_bd_golden.py from the reference itself) ->
    [(seed, chromosome, window, events, queries, clusters per query)]"""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bd_clusters.npz"))
    out = []
    for seed in z["seeds"]:
        chr_id, ws, we = (int(v) for v in z["win_%d" % seed])
        cnt, reg = z["cnt_%d" % seed], z["reg_%d" % seed]
        off = np.concatenate([[0], np.cumsum(cnt)])
        cl = [[(int(c), int(s), int(e)) for c, s, e in reg[off[i]:off[i + 1]]] for i in range(cnt.size)]
        out.append((int(seed), chr_id, (ws, we), z["ev_%d" % seed].astype(np.int64), z["q_%d" % seed], cl))
    return out
