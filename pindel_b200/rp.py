"""Host side of the discordant-read-pair clustering (include/pindel_pg.h: pg_rp_modify, pg_rp_summarize): the record
layouts, and the order ModifyRP's sort produces for distinct keys (src/bddata.cpp:239-253, Compare2RP: OriginalPosA
descending, then OriginalPosB descending -- among equal keys std::sort decides, which is why the C ABI takes sorted rows)."""
from __future__ import annotations

import numpy as np

RP_DT = np.dtype([("pos_a", "<u4"), ("pos_b", "<u4"), ("insert_size", "<u4"), ("read_length", "<i2"), ("da", "u1"), ("db", "u1")])
CLUSTER_DT = np.dtype([("pos_a", "<u4"), ("pos_a1", "<u4"), ("pos_b", "<u4"), ("pos_b1", "<u4"), ("n_identical", "<u4"),
                       ("merged_into", "<u4"), ("visited", "u1"), ("report", "u1"), ("pad", "<u2")])
assert RP_DT.itemsize == 16 and CLUSTER_DT.itemsize == 28


def compare2rp_order(rows: np.ndarray) -> np.ndarray:
    """A permutation that satisfies Compare2RP (ties in input order)."""
    return np.lexsort((-rows["pos_b"].astype(np.int64), -rows["pos_a"].astype(np.int64)))
