"""Development tool: one search launch on the 3.08 Gbp / 24-contig workload of bench.py (the only shape whose packed reference,
1.16 GB, does not fit the 126 MB L2), for an ncu capture of its DRAM traffic:

    ncu --set full --clock-control none --import-source on -k regex:pg_search_kernel -c 1 -o gpurun_out/wg python scripts/wg_probe.py [reads]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from pindel_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
a = argparse.Namespace(wg_reads=n)
spec = bench.spec_wg(a, 1)
t = time.time()
w = bench.make_workload(spec, 0, 1, torch.device("cuda", 0))
print("workload in %.1fs" % (time.time() - t), flush=True)
pg = api.PatternGrowth()
pg.set_reference(w["padded"])
L = w["seq"].shape[1]
pg.upload(seq_blob=w["seq"].reshape(-1), seq_off=(np.arange(n + 1, dtype=np.uint64) * L).astype(np.uint32), matched_d=w["md"], matched_rel_pos=w["pos"],
          insert_size=w["isz"], chr_id=w["chr_id"])
for _ in range(3):
    t = time.perf_counter()
    pg.launch(3)
    pg.sync()
    dt = time.perf_counter() - t
    print("%d reads on %d contigs: %.2f M reads/s" % (n, len(w["padded"]), n / dt / 1e6), flush=True)
pg.close()
