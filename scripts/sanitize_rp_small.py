"""racecheck-sized workload for the shared-memory paths that scripts/sanitize_gpu.py is too large for under racecheck:
one CTA-sized read-pair class (pg_rp_modify_kernel) and a few BGZF blocks (pg_bgzf_inflate_kernel)."""
import sys
sys.path.insert(0, ".")
from pindel_b200 import api, bgzf, rp, synth
from oracle import rporacle

pg = api.PatternGrowth()
rows = synth.make_discordant_pairs(3, 4, span=20_000, max_support=900)
rows = rows[rp.compare2rp_order(rows)]
mod = pg.rp_modify(rows)
assert mod.tobytes() == rporacle.modify(rows).tobytes()
print("rp ok", pg.rp_last_classes())
payload = synth.make_bam_like_bytes(150_000, 9)
assert pg.bgzf_inflate(bgzf.compress(payload)).tobytes() == payload
print("bgzf ok")
pg.close()
