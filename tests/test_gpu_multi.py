"""The N > 1 path on real GPUs: two ranks (NCCL), each searches its shard of one batch on its own B200 through the C ABI,
the breakpoint records go to rank 0 with the gatherer bench.py uses (straight from the library's result slots, on a side
stream, while the next launch is already queued), are merged back into batch order and compared with the oracle.  Skips on
boxes with fewer than two GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
from pindel_b200 import api, shard, synth, runs as R
from tests import helpers as H
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world, device_id=dev)
ref = synth.make_reference(1_500_000, 3001)
padded = synth.pad_reference(ref)
batch = synth.make_reads(ref, 6000, 100, seed=3002)
owners = shard.owner_of_reads(batch.matched_rel_pos, None, world, shard=100_000)
assert set(owners.tolist()) == {0, 1}
mine = np.nonzero(owners == rank)[0]
sub = synth.ReadBatch(batch.seq[mine], batch.matched_d[mine], batch.matched_rel_pos[mine], batch.insert_size[mine], batch.kind[mine])
pg = api.PatternGrowth(device=rank, max_range_index=%(x)d)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
pg.set_stream(stream.cuda_stream)
pg.set_reference([padded])
pg.upload(**H.batch_kwargs(sub))
gat = shard.RecordGatherer(dev, rank, world)
n = sub.n
got = []
# three launches back to back; the records of launch i are gathered while launch i+1 is queued
evs = []
for i in range(3):
    pg.launch(3)
    ev = torch.cuda.Event(); ev.record(stream); evs.append(ev)
    if i > 0:
        o, r, nr, nn = pg.device_results_previous()
        assert nn == n
        gat.gather(o, r, nn * R.OUT_DT.itemsize, nr * R.RUN_DT.itemsize, after_event=evs[i - 1], to_host=(i == 2))
        if rank == 0:
            gat.stream.synchronize()
            got.append([(a.cpu().numpy().view(R.OUT_DT).copy(), b.cpu().numpy().view(R.RUN_DT).copy()) for a, b in gat.last])
o, r, nr = pg.device_results()
gat.gather(o, r, n * R.OUT_DT.itemsize, nr * R.RUN_DT.itemsize, after_event=evs[-1])
gat.stream.synchronize()
if rank == 0:
    got.append([(a.cpu().numpy().view(R.OUT_DT).copy(), b.cpu().numpy().view(R.RUN_DT).copy()) for a, b in gat.last])
    fb, fo = batch.blob()
    exp = H.oracle_results([padded], batch, threads=8, max_range_index=%(x)d)
    for k, parts in enumerate(got):
        mo, mr = shard.merge_gathered(parts, owners)
        H.compare(exp, R.per_read(mo, mr, fb, fo), "2-GPU run, gather %%d" %% k)
    print("NCCL-OK", len(exp), len(got))
dist.barrier()
pg.close()
dist.destroy_process_group()
"""


@pytest.mark.parametrize("x", [2, 4])
def test_two_gpus_shard_gather_merge(x):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29500 + (os.getpid() % 2000) + x
    code = WORKER % dict(root=ROOT, port=port, x=x)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=900) for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[0][-1500:] + o[1][-1500:] for o in outs)
    assert "NCCL-OK 6000 3" in outs[0][0]
