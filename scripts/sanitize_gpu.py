"""Small GPU workload for compute-sanitizer (memcheck / racecheck / synccheck): edge cases + a few hundred synthetic reads."""
import sys
sys.path.insert(0, ".")
import numpy as np
from pindel_b200 import api, runs as R, synth
from tests import fixtures as F, helpers as H

case = F.load_edge()
pg = api.PatternGrowth()
pg.set_reference([case.padded])
kw = case.kwargs()
out, runs = pg.search(**kw)
H.compare(case.expected, R.per_read(out, runs, kw["seq_blob"], kw["seq_off"]), "sanitizer edge")
ref = synth.make_reference(300_000, 5, n_runs=[(100_000, 400)])
b = synth.make_reads(ref, 600, 150, seed=6)
pg.set_reference([synth.pad_reference(ref)])
out, runs = pg.search(**H.batch_kwargs(b))
print("sanitizer workload ok", out.size, runs.size)
pg.close()

# rows N1-N3: classification (host arrays and resident), read-pair clustering (warp and CTA classes), BGZF inflation
from pindel_b200 import bgzf, rp, sv
pg = api.PatternGrowth()
padded = synth.pad_reference(ref)
pg.set_reference([padded])
out, runs = pg.search(**H.batch_kwargs(b))
win = sv.window(int(padded.size), int(ref.size), region=(0, int(ref.size)))
rec = pg.classify_resident(win)
keep = np.nonzero(out["n_close_runs"] > 0)[0]
print("classify_resident ok", rec.size, "records for", keep.size, "reads with a close end")
rows = synth.make_discordant_pairs(60, 3, span=3_000_000, background=60)
rows = rows[rp.compare2rp_order(rows)]
fin = pg.rp_summarize(rows, pg.rp_modify(rows))
dense = synth.make_discordant_pairs(40, 4, span=200_000, max_support=200)       # one CTA class per strand combination
dense = dense[rp.compare2rp_order(dense)]
fin2 = pg.rp_summarize(dense, pg.rp_modify(dense))
print("rp ok", int(fin["report"].sum()), int(fin2["report"].sum()), pg.rp_last_classes())
payload = synth.make_bam_like_bytes(400_000, 9)
assert pg.bgzf_inflate(bgzf.compress(payload)).tobytes() == payload
assert pg.bgzf_inflate(bgzf.compress(payload, level=0, block=0x8000)).tobytes() == payload
print("bgzf ok")
pg.close()
