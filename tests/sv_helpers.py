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
