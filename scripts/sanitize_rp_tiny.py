"""racecheck-sized workload for pg_rp_modify_kernel (the CTA-wide recurrence with the class's intervals in shared memory):
classes of a few dozen to ~250 rows forced onto the CTA kernel with PG_RP_WARP_MAX=8, result compared with the oracle.

    PG_RP_WARP_MAX=8 compute-sanitizer --tool racecheck python scripts/sanitize_rp_tiny.py"""
import os
import sys
sys.path.insert(0, ".")
os.environ.setdefault("PG_RP_WARP_MAX", "8")
from pindel_b200 import api, rp, synth
from oracle import rporacle

pg = api.PatternGrowth()
rows = synth.make_discordant_pairs(4, 11, span=30_000, max_support=120)
rows = rows[rp.compare2rp_order(rows)]
mod = pg.rp_modify(rows)
assert mod.tobytes() == rporacle.modify(rows).tobytes()
cls = pg.rp_last_classes()
assert cls[1] >= 1, cls
print("rp ok: %d rows, (warp classes, CTA classes, rows in classes) = %s" % (rows.size, (cls,)))
pg.close()
