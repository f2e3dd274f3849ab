"""The N > 1 path on CPU: two processes (gloo), each searches its shard of the batch -- with the emulated warp
program standing in for the GPU kernel -- then one gather on rank 0 and a merge back into batch order.  The merged
result must equal the single-process oracle result."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
from pindel_b200 import shard, synth, runs as R
from tests.emul import emul
from tests import helpers as H
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world)
ref = synth.make_reference(900_000, 3001)
padded = synth.pad_reference(ref)
batch = synth.make_reads(ref, 240, 100, seed=3002)
owners = shard.owner_of_reads(batch.matched_rel_pos, None, world, shard=100_000)
assert set(owners.tolist()) == {0, 1}
mine = np.nonzero(owners == rank)[0]
sub = synth.ReadBatch(batch.seq[mine], batch.matched_d[mine], batch.matched_rel_pos[mine], batch.insert_size[mine], batch.kind[mine])
lib = emul.build()
prm = emul.make_params(lib)
blob, off = sub.blob()
out, pool = emul.search(lib, prm, [padded], blob, off, sub.matched_d, sub.matched_rel_pos, sub.insert_size)
parts = shard.gather_records(torch.from_numpy(out.view(np.uint8).copy()), torch.from_numpy(pool.view(np.uint8).copy()), dst=0)
if rank == 0:
    parts = [(o.numpy().view(R.OUT_DT), r.numpy().view(R.RUN_DT)) for o, r in parts]
    mo, mr = shard.merge_gathered(parts, owners)
    fb, fo = batch.blob()
    got = R.per_read(mo, mr, fb, fo)
    exp = H.oracle_results([padded], batch, threads=2)
    H.compare(exp, got, "2-rank gloo run")
    print("GLOO-OK", len(got))
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_shard_gather_merge():
    from tests.emul import emul
    emul.build()       # build once, before the ranks race for it
    port = 29000 + (os.getpid() % 2000)
    code = WORKER % dict(root=ROOT, port=port)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[0][-1500:] + o[1][-1500:] for o in outs)
    assert "GLOO-OK 240" in outs[0][0]


def test_shard_windows_cover_and_partition():
    from pindel_b200 import shard
    for world in (1, 2, 4, 8):
        wins = sorted(set(w for r in range(world) for w in shard.shard_windows(62_435_964, r, world)))
        if world == 1:
            assert wins == [(0, 62_435_964)]
        else:
            assert wins[0][0] == 0 and wins[-1][1] == 62_435_964
            assert all(a[1] == b[0] for a, b in zip(wins, wins[1:]))
