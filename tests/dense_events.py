"""Seeded Pindel-text input that is dense in the events whose reporters order their boxes with inlined loops: deletions with
non-template bases (SortOutputDI) and inversions with and without non-template bases (OutputSorter), plus plain deletions,
short insertions and tandem duplications for the other three reporters, many reads per event
from both anchor strands, a handful of breakpoint offsets (ties in every key, the same fragment seen several times), events
a few bases apart (several events per box).  Test input only: both programs must write the same files from it."""
import numpy as np

COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    COMP[_a] = _b
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _rc(x):
    return COMP[x[::-1]]


def _model(ref, kind, a, k, sz, nt, L):
    """the read as it lies on the strand `ref` is given in: k bases up to the breakpoint a, then the event's other side"""
    first = ref[a - k:a]
    rest = L - k - len(nt)
    if kind in ("D", "DI"):
        second = ref[a + sz:a + sz + rest]
    elif kind == "SI":                      # short insertion: the non-template bases, then straight on
        second = ref[a:a + rest]
    elif kind == "TD":                      # tandem duplication: the unit's end followed by its start
        second = ref[a - sz:a - sz + rest]
    else:                                   # inversion: the far side read backwards, complemented
        second = COMP[ref[a + sz - rest:a + sz][::-1]]
    return np.concatenate([first, nt, second])


def make(seed=11, G=400_000, L=100, insert=500):
    """-> (reference uint8 [G], Pindel text, number of reads)"""
    rng = np.random.default_rng(seed)
    ref = ACGT[rng.integers(0, 4, G)]
    ref_rc = _rc(ref)
    events = []
    pos = 20_000
    while pos < G - 30_000:
        kind = str(rng.choice(["DI", "DI", "INV", "INV", "INV_NT", "D", "SI", "TD"]))
        sz = int(rng.integers(5, 300)) if kind in ("D", "DI") else (0 if kind == "SI" else int(rng.integers(120, 1500)))
        nt = ACGT[rng.integers(0, 4, int(rng.integers(1, 9)))] if kind in ("DI", "INV_NT", "SI") else ACGT[:0]
        events.append((kind, pos, sz, nt))
        if rng.random() < 0.4:              # a second event a few bases on: same box, other breakpoints
            events.append((kind, pos + int(rng.integers(2, 30)), max(0, sz + int(rng.integers(-3, 4))) if kind != "SI" else 0,
                           nt if rng.random() < 0.5 else ACGT[rng.integers(0, 4, len(nt))]))
        pos += int(rng.integers(3000, 6000))
    lines, n = [], 0
    for kind, a, sz, nt in events:
        ks = rng.integers(30, L - 30, 6)                       # few offsets: ties in BP, fragments seen more than once
        for _ in range(int(rng.integers(8, 40))):
            k = int(rng.choice(ks))
            plus = rng.random() < 0.5
            if plus:
                t = _model(ref, kind, a, k, sz, nt, L)
                g_first = a - k
            else:                                              # the same event from the other strand
                a2 = G - a if kind == "SI" else (G - (a - sz) if kind == "TD" else G - (a + sz))
                k = L - k - len(nt)
                t = _model(ref_rc, kind, a2, k, sz, _rc(nt), L)
                g_first = a2 - k
            if len(t) != L:
                continue
            t = t.copy()
            err = rng.random(L) < 0.003
            t[err] = ACGT[rng.integers(0, 4, int(err.sum()))]
            slack = int(rng.choice([40, 80, 120, 200]))
            mpos = g_first - slack if plus else (G - 1 - g_first) + slack
            lines.append("@d%d" % n)
            lines.append(_rc(t).tobytes().decode())
            lines.append("%s\tdense\t%d\t60\t%d\tS" % ("+" if plus else "-", mpos, insert))
            n += 1
    return ref, "\n".join(lines) + "\n", n
