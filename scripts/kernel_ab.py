"""Development tool: kernel-only timing of several builds of libpindel_pg.so on the headline workload (code layout and
register allocation move this kernel's speed by several percent, so variants are compared on the hardware, side by side).

    python scripts/kernel_ab.py build/variants/libpg_a.so build/variants/libpg_b.so ...      (on the GPU box)

Each library is loaded in its own process (PG_LIB); the reads are generated once (torch, on the GPU) and shared through
/dev/shm.  Prints reads/s of back-to-back launches at -x 2 (4 M reads) and -x 5 (1 M reads)."""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHM = "/dev/shm/pg_ab"

if len(sys.argv) > 1 and sys.argv[1] == "--worker":
    from pindel_b200 import api, synth
    z = {k: np.load(os.path.join(SHM, k + ".npy")) for k in ("padded", "seq", "md", "pos", "isz")}
    n, L = z["seq"].shape
    out = []
    for x, m in ((2, n), (5, n // 4)):
        pg = api.PatternGrowth(max_range_index=x)
        pg.set_reference([z["padded"]])
        pg.upload(seq_blob=z["seq"][:m].reshape(-1), seq_off=(np.arange(m + 1, dtype=np.uint64) * L).astype(np.uint32), matched_d=z["md"][:m],
                  matched_rel_pos=z["pos"][:m], insert_size=z["isz"][:m])
        for _ in range(2):
            pg.launch(3)
        pg.sync()
        k = 4
        t = time.perf_counter()
        for _ in range(k):
            pg.launch(3)
        pg.sync()
        dt = (time.perf_counter() - t) / k
        out.append("-x %d: %7.2f M reads/s" % (x, m / dt / 1e6))
        pg.close()
    print("%-40s %s" % (os.path.basename(os.environ.get("PG_LIB", "default")), "   ".join(out)), flush=True)
    sys.exit(0)

import torch
from pindel_b200 import synth
os.makedirs(SHM, exist_ok=True)
ref = synth.make_reference(62_435_964, 20260926)
rt = torch.from_numpy(ref).cuda()
seq, md, pos, isz, _ = synth.make_reads_torch(rt, 4_000_000, 150, 777, insert_size=500)
np.save(os.path.join(SHM, "padded.npy"), synth.pad_reference(ref))
np.save(os.path.join(SHM, "seq.npy"), seq.cpu().numpy())
np.save(os.path.join(SHM, "md.npy"), md.cpu().numpy())
np.save(os.path.join(SHM, "pos.npy"), pos.cpu().numpy().astype(np.uint32))
np.save(os.path.join(SHM, "isz.npy"), isz.cpu().numpy())
del rt, seq
torch.cuda.empty_cache()
for rep in range(int(os.environ.get("AB_REPS", "2"))):
    for lib in sys.argv[1:]:
        subprocess.run([sys.executable, os.path.abspath(__file__), "--worker"], env=dict(os.environ, PG_LIB=os.path.abspath(lib)))
