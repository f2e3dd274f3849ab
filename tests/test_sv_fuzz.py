"""Adversarial inputs for the SV-classification program (pg_sv.cuh compiled for the host) against the driver oracle
(oracle/sv_oracle.c, itself pinned to the reference's drivers by test_sv_oracle_pinning.py).  The lists here are not what
a search produces: several close runs with one terminal, many far runs on either strand, mismatch counts in any order,
far ends placed so that many pairs pass an arm's test, regions / box counts / window ends chosen so that reads are
*not* taken at the first fire -- the cases where the order of fires and the last fire decide the result.  No GPU."""
import numpy as np
import pytest

from oracle import pgoracle, svoracle
from pindel_b200 import sv, synth
from tests import sv_helpers as S
from tests.emul import sv_emul

SP = 100000


def _close_lists(rng, rl, anchor):
    """1-3 runs sharing direction, strand and terminal (+ sometimes a stray run CleanUniquePoints would drop)."""
    d = int(rng.integers(0, 2))
    st = int(rng.integers(0, 2))
    term = anchor
    rows = []
    ln = int(rng.integers(5, 12))
    for _ in range(int(rng.integers(1, 4))):
        cnt = int(rng.integers(1, max(2, rl // 2)))
        mm = int(rng.integers(0, 4))
        for k in range(cnt):
            if ln >= rl:
                break
            rows.append([ln, term + ln if d == 0 else term - ln, d, st, mm])
            ln += 1
        ln += int(rng.integers(0, 3))
    if rng.random() < 0.2 and len(rows) > 1:        # a stray point with another terminal, not at the end
        rows.insert(int(rng.integers(0, len(rows) - 1)), [7, term + 50, d, st, 0])
    return np.array(rows, dtype=np.int64).reshape(-1, 5)


def _far_lists(rng, rl, close):
    rows = []
    c_loc = close[:, 1]
    for _ in range(int(rng.integers(1, 5))):
        d = int(rng.integers(0, 2))
        st = int(rng.integers(0, 2))
        mm = int(rng.integers(0, 4))
        mode = rng.random()
        ln = int(rng.integers(8, max(9, rl - 8)))
        cnt = int(rng.integers(1, max(2, rl - ln)))
        if mode < 0.35:      # adjacent to a close point (short insertions)
            loc = int(c_loc[int(rng.integers(0, len(c_loc)))]) + int(rng.integers(-2, 3))
        elif mode < 0.8:     # within a few hundred bases either side (deletions, duplications, small inversions)
            loc = int(c_loc[-1]) + int(rng.integers(-400, 400))
        else:
            loc = int(c_loc[-1]) + int(rng.integers(-30000, 30000))
        for k in range(cnt):
            rows.append([ln + k, loc + k if d == 0 else loc - k, d, st, mm])
    return np.array(rows, dtype=np.int64).reshape(-1, 5)


@pytest.mark.parametrize("seed,region_kind,boxes_kind", [(1, "all", "all"), (2, "narrow", "all"), (3, "all", "few"), (4, "narrow", "few"),
                                                         (5, "all", "all"), (6, "half", "few")])
def test_adversarial_lists(seed, region_kind, boxes_kind):
    rng = np.random.default_rng(1000 + seed)
    nbases = 200_000
    ref = synth.make_reference(nbases, 4000 + seed, n_runs=[(60_000, 500)])
    # low-complexity stretches so that the breakpoint slides (GetRealStart4*, LeftMostTD/INV) actually move
    for s in range(5000, nbases - 5000, 7000):
        ref[s:s + 200] = np.frombuffer(b"ACACACACAT" * 20, dtype=np.uint8)
    padded = synth.pad_reference(ref)
    n = 1500
    seqs, md, isz, closes, fars = [], [], [], [], []
    for i in range(n):
        rl = int(rng.integers(30, 130))
        anchor = SP + int(rng.integers(2000, nbases - 2000))
        if rng.random() < 0.5:   # read text taken from the reference so that inserted strings match their surroundings
            seqs.append(bytes(ref[anchor - SP:anchor - SP + rl]))
        else:
            seqs.append(bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), rl, p=[0.245, 0.245, 0.245, 0.245, 0.02])))
        md.append(ord("+-"[int(rng.integers(0, 2))]))
        isz.append(int(rng.integers(100, 600)))
        c = _close_lists(rng, rl, anchor)
        closes.append(c)
        fars.append(_far_lists(rng, rl, c) if rng.random() < 0.95 else np.zeros((0, 5), np.int64))
    md = np.array(md, dtype=np.uint8)
    isz = np.array(isz, dtype=np.int16)
    rflags = (rng.random(n) < 0.05).astype(np.uint8) | ((rng.random(n) < 0.03).astype(np.uint8) << 1)
    region = {"all": (0, nbases), "narrow": (90_000, 110_000), "half": (0, nbases // 2)}[region_kind]
    win = sv.window(int(padded.size), nbases if seed % 2 else nbases // 2, region=region,
                    num_boxes=None if boxes_kind == "all" else 12_000)
    # oracle side: expanded rows, close lists as given (the oracle's drivers see the cleaned list)
    cleaned = []
    for c in closes:
        last = c[-1]
        term = np.where(c[:, 2] == 0, c[:, 1] - c[:, 0], c[:, 1] + c[:, 0])
        lt = last[1] - last[0] if last[2] == 0 else last[1] + last[0]
        cleaned.append(c[(term == lt) & (c[:, 2] == last[2]) & (c[:, 3] == last[3])])
    rows = np.concatenate([np.concatenate((c, f)) for c, f in zip(cleaned, fars)])
    up_off = np.zeros(n + 1, dtype=np.uint32)
    up_off[1:] = np.cumsum([len(c) + len(f) for c, f in zip(cleaned, fars)])
    blob = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    off = np.zeros(n + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(s) for s in seqs])
    P = svoracle.make_params(pgoracle.prob_table(), win["window_end"], win["num_boxes"], win["box_size"],
                             region=((1, region[0]), (1, region[1])))
    exp = svoracle.classify(P, padded.tobytes(), blob, off, md, isz, rflags, (rows & 0xFFFFFFFF).astype(np.uint32).view(np.int32),
                            up_off, np.array([len(c) for c in cleaned], np.uint32))
    # product side: raw close lists, run-length coded
    b = sv.batch_from_lists(seqs, md, isz, closes, fars, rflags=rflags)
    rec = sv_emul.classify(sv_emul.build(), padded, b, win)
    got = S.dump_from_records(rec, seqs, rflags)
    svoracle.compare_dumps(exp, got, "fuzz seed %d" % seed)
    # the generator does produce reads the drivers take, and reads whose fields were written without the read being taken
    taken = sum(int(exp[d]["vars"].shape[0]) + int(exp[d]["future"].shape[0]) for d in exp)
    last = exp[max(exp)]["state"]
    touched_not_used = int(((last[:, 7] == 0) & (last[:, :6] != 0).any(axis=1)).sum())
    assert taken > 20, taken
    if region_kind != "all" or boxes_kind != "all":
        assert touched_not_used > 5, touched_not_used
