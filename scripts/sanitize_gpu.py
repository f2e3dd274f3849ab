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
