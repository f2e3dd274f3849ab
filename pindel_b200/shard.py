"""Multi-GPU layout of the search path: reads are independent units, so the batch is sharded by anchor window and
the only collective is one gather of the breakpoint records (per-read headers + runs) on rank 0, where the
reference's drivers/reporter would run once (windows are the reference's own unit of work,
src/pindel.cpp:1778-1985).  Works with any torch.distributed backend: NCCL on the GPUs, gloo in the CPU tests."""
from __future__ import annotations

import numpy as np

from . import runs as R

SHARD_BASES = 10_000_000      # BASELINE.json configs[3]: 10 Mbp anchor-window shards


def shard_windows(ref_bases: int, rank: int, world: int, shard: int = SHARD_BASES):
    """Windows [lo, hi) of one chromosome owned by `rank`: dealt round-robin."""
    if world == 1:
        return [(0, ref_bases)]
    wins = [(s, min(s + shard, ref_bases)) for s in range(0, ref_bases, shard)]
    mine = [w for i, w in enumerate(wins) if i % world == rank]
    return mine or [wins[rank % len(wins)]]


def owner_of_reads(matched_rel_pos: np.ndarray, chr_id, world: int, shard: int = SHARD_BASES) -> np.ndarray:
    """Rank that searches each read: its anchor's window index (and chromosome) modulo the world size."""
    w = (np.asarray(matched_rel_pos, dtype=np.int64) // shard)
    if chr_id is not None:
        w = w + np.asarray(chr_id, dtype=np.int64) * 7919
    return (w % world).astype(np.int32)


def gather_records(out_t, runs_t, dst: int = 0):
    """One gather of (pg_read_out[n_local], pg_run[n_runs_local]) byte tensors to rank `dst`.

    out_t / runs_t: uint8 torch tensors (device of the backend).  Sizes differ per rank, so the sizes travel first
    (a 2-element all_gather) and the payload is padded to the maximum.  Returns, on dst, a list of
    (out_bytes, runs_bytes) per rank (trimmed); None elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    rank = dist.get_rank()
    sz = torch.tensor([out_t.numel(), runs_t.numel()], dtype=torch.int64, device=out_t.device)
    sizes = [torch.zeros_like(sz) for _ in range(world)]
    dist.all_gather(sizes, sz)
    sizes = torch.stack(sizes).cpu().numpy()
    cap = int(sizes[:, 0].max() + sizes[:, 1].max())
    payload = torch.zeros(cap, dtype=torch.uint8, device=out_t.device)
    payload[:out_t.numel()] = out_t
    payload[int(sizes[:, 0].max()):int(sizes[:, 0].max()) + runs_t.numel()] = runs_t
    if rank == dst:
        bufs = [torch.empty_like(payload) for _ in range(world)]
        dist.gather(payload, bufs, dst=dst)
        o0 = int(sizes[:, 0].max())
        return [(bufs[r][:int(sizes[r, 0])], bufs[r][o0:o0 + int(sizes[r, 1])]) for r in range(world)]
    dist.gather(payload, None, dst=dst)
    return None


def merge_gathered(parts, owners: np.ndarray):
    """Rank 0: stitch the per-rank records back into the original read order.
    parts: list over ranks of (out ndarray OUT_DT, runs ndarray RUN_DT); owners[i] = rank that searched read i
    (reads keep their relative order inside a shard).  Returns (out, runs) for the whole batch."""
    n = owners.size
    out = np.zeros(n, dtype=R.OUT_DT)
    base = 0
    runs_all = []
    for r, (o, ru) in enumerate(parts):
        idx = np.nonzero(owners == r)[0]
        assert idx.size == o.size, (r, idx.size, o.size)
        oo = o.copy()
        oo["run_off"] = oo["run_off"] + base
        out[idx] = oo
        runs_all.append(ru)
        base += ru.size
    return out, (np.concatenate(runs_all) if runs_all else np.zeros(0, dtype=R.RUN_DT))
