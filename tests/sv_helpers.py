"""Helpers of the SV-classification tests: turn per-read search results into the classify inputs, window parameters."""
from __future__ import annotations

import numpy as np

from pindel_b200 import runs as R


def final_seqs(blob, off, per_read):
    """UnmatchedSeq of every read after GetCloseEnd (plain A/C/G/T/N reads: a flip is one reverse complement)."""
    out = []
    for i, r in enumerate(per_read):
        s = R.set_unmatched_seq(bytes(blob[int(off[i]):int(off[i + 1])]))
        if r["flip"]:
            s = R.set_unmatched_seq(R.reverse_complement(s))
        assert len(s) == r["L"], (i, len(s), r["L"])
        out.append(s)
    return out


def classify_inputs(per_read, seqs, matched_d, insert_size, keep_only_close=True):
    """per_read: dicts with `close` (cleaned) and `far` UP rows [k,6].  Reads without a close end are dropped, as the
    reference's readers do (src/reader.cpp:262-288).  -> dict of arrays + `index` (positions in the original batch)."""
    idx = [i for i, r in enumerate(per_read) if len(r["close"]) or not keep_only_close]
    rows, up_off, n_close = [], [0], []
    for i in idx:
        c, f = per_read[i]["close"], per_read[i]["far"]
        rows.append(c[:, :5])
        rows.append(f[:, :5])
        n_close.append(len(c))
        up_off.append(up_off[-1] + len(c) + len(f))
    up = np.concatenate(rows).astype(np.int64) if rows else np.zeros((0, 5), np.int64)
    up = (up & 0xFFFFFFFF).astype(np.uint32).view(np.int32).reshape(-1, 5)
    blob = np.frombuffer(b"".join(seqs[i] for i in idx), dtype=np.uint8)
    off = np.zeros(len(idx) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(seqs[i]) for i in idx])
    return dict(index=np.array(idx), seq_blob=blob, seq_off=off, matched_d=np.asarray(matched_d)[idx].astype(np.uint8),
                insert_size=np.asarray(insert_size)[idx].astype(np.int16), rflags=np.zeros(len(idx), np.uint8),
                up_rows=up, up_off=np.array(up_off, np.uint32), n_close=np.array(n_close, np.uint32))


def window_params(comp_size, window_end=None):
    """BoxSize / NumBoxes as main() derives them (src/pindel.cpp:1806-1810)."""
    box = comp_size // 30000 or 1
    return dict(box_size=box, num_boxes=comp_size * 2 // box + 1, window_end=comp_size if window_end is None else window_end)


def _fnv(s: bytes) -> int:
    h = 2166136261
    for c in s:
        h = ((h ^ c) * 16777619) & 0xFFFFFFFF
    return h


def _state_row(s):
    nt = s.NT_str
    head = int.from_bytes(nt[:4].ljust(4, b"\0"), "little")
    tail = int.from_bytes(nt[::-1][:4].ljust(4, b"\0"), "little")
    v = [s.Left, s.Right, s.BP, s.BPLeft, s.BPRight, s.IndelSize, s.NT_size, int(s.Used), len(nt), _fnv(nt), head, tail]
    return np.array(v, dtype=np.int64).astype(np.uint32).view(np.int32) if False else [(x + (1 << 32)) % (1 << 32) for x in v]


def dump_from_records(records, seqs, rflags, driver_mask=0x7F):
    """Replay device records (pindel_b200.sv.replay) into the dump layout of oracle/sv_oracle.c, so that product output and
    oracle output compare with oracle.svoracle.compare_dumps."""
    from pindel_b200 import sv
    n = len(seqs)
    _, per = sv.replay(records, seqs, rflags)
    out = {}
    last_state = [sv.ReadState(bool(f & 1)) for f in rflags]
    for d in range(7):
        if not (driver_mask >> d) & 1:
            continue
        if d in per:
            p = per[d]
            last_state = p["state"]
            v = np.array(p["vars"], dtype=np.int64).reshape(-1, 2)
            f = np.array([[i] + _state_row(s) for i, s in p["future"]], dtype=np.int64).reshape(-1, 13)
        else:
            v, f = np.zeros((0, 2), np.int64), np.zeros((0, 13), np.int64)
        st = np.array([_state_row(s) for s in last_state], dtype=np.int64).reshape(n, 12)
        to32 = lambda a: (a & 0xFFFFFFFF).astype(np.uint32).view(np.int32)
        out[d] = dict(vars=to32(v), future=to32(f), state=to32(st))
    return out


def load_sv_case(name):
    """A committed SV fixture (tests/golden/sv_<name>.npz, made by make_sv_golden.py from the reference itself):
    -> dict(case, ci (classify inputs, expanded rows), seqs (final sequences of the kept reads), z (npz), golden (parsed dump),
            win (pindel_b200.sv.window dict), oracle_params)."""
    import os
    from oracle import svoracle
    from pindel_b200 import sv
    from tests.golden import make_sv_golden as M
    case, ci, _, _, seqs_all = M.case_inputs(name)
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sv_%s.npz" % name))
    n = int(z["n"])
    assert n == len(ci["index"])
    region = (int(z["region"][0]), int(z["region"][1]))
    win = sv.window(int(case.padded.size), int(z["window_end"]), driver_mask=int(z["driver_mask"]), region=region,
                    min_inversion_size=int(z["min_inv"]), min_matched_bases=int(z["min_matched"]), box_size=int(z["box_size"]),
                    num_boxes=int(z["num_boxes"]))
    P = svoracle.make_params(z["table"], win["window_end"], win["num_boxes"], win["box_size"], driver_mask=win["driver_mask"],
                             region=((1, region[0]), (1, region[1])), min_inv=win["min_inversion_size"],
                             min_matched=win["min_matched_bases"], seq_error_rate=float(z["seq_error_rate"]))
    return dict(case=case, ci=ci, seqs=[seqs_all[i] for i in ci["index"]], z=z, golden=svoracle.parse_dump(z["stream"], n), win=win,
                oracle_params=P, seq_error_rate=float(z["seq_error_rate"]), sensitivity=float(case.settings.get("sensitivity", 0.95)))


def oracle_dump(c):
    from oracle import svoracle
    kw = {k: c["ci"][k] for k in ("seq_blob", "seq_off", "matched_d", "insert_size", "rflags", "up_rows", "up_off", "n_close")}
    return svoracle.classify(c["oracle_params"], c["case"].padded.tobytes(), **kw)


def product_batch(c, raw_close=False):
    """The pg_sv_reads arrays of a fixture (UP lists run-length coded)."""
    from pindel_b200 import sv
    idx = c["ci"]["index"]
    exp = c["case"].expected
    return sv.batch_from_lists(c["seqs"], c["ci"]["matched_d"], c["ci"]["insert_size"],
                               [exp[i]["close_raw" if raw_close else "close"] for i in idx], [exp[i]["far"] for i in idx],
                               rflags=c["ci"]["rflags"])
