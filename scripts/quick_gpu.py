"""Ad-hoc GPU smoke + timing used during development: python scripts/quick_gpu.py [n_reads] [L] [ref_mbp]."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from pindel_b200 import api, synth, runs as R

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 150
mbp = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ref = synth.make_reference(mbp * 1_000_000, 20260926)
padded = synth.pad_reference(ref)
t = time.time(); batch = synth.make_reads(ref, n, L, seed=5); print("gen %.2fs" % (time.time() - t))
import os
pg = api.PatternGrowth(max_range_index=int(os.environ.get('PG_X', '2')))
t = time.time(); pg.set_reference([padded]); print("set_reference %.3fs" % (time.time() - t))
blob, off = batch.blob()
kw = dict(seq_blob=blob, seq_off=off, matched_d=batch.matched_d, matched_rel_pos=batch.matched_rel_pos, insert_size=batch.insert_size)
for it in range(3):
    t = time.time(); out, runs = pg.search(**kw); dt = time.time() - t
    print("search e2e %.4fs  %.3f Mreads/s  runs=%d" % (dt, n / dt / 1e6, runs.size))
pg.upload(**kw)
for it in range(3):
    t = time.time(); pg.launch(3); pg.sync(); dt = time.time() - t
    print("kernel only %.4fs  %.3f Mreads/s" % (dt, n / dt / 1e6))
if len(sys.argv) > 4:      # back-to-back launches, like bench.py's timed region
    k = int(sys.argv[4])
    t = time.time()
    for it in range(k):
        pg.launch(3)
    pg.sync(); dt = (time.time() - t) / k
    print("kernel x%d back to back %.4fs  %.3f Mreads/s" % (k, dt, n / dt / 1e6))
res = R.per_read(out[:2000], runs)
print("close found", sum(1 for r in res if len(r["close"])), "far found", sum(1 for r in res if len(r["far"])), "of 2000")
