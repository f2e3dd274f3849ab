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


class DevArr:
    """A raw device pointer as a __cuda_array_interface__ object (for torch.as_tensor)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class RecordGatherer:
    """The single collective of the path on the GPUs: every rank's breakpoint records (per-read headers + runs, straight from
    the library's result slot: pg_device_results / pg_device_results_previous) to rank 0 over NCCL, on its own stream so that
    it travels while the next launch runs.  Exact sizes (one small all_gather of the two byte counts), no padding, no staging
    copy; rank 0 keeps one receive buffer per rank and, when asked, copies everything on to pinned host memory."""

    def __init__(self, dev, rank, world):
        import torch
        self.dev, self.rank, self.world = dev, rank, world
        self.stream = torch.cuda.Stream(device=dev)
        self.bufs = [None] * world
        self.host = None
        self.last = None            # rank 0: [(out bytes tensor, runs bytes tensor)] per rank, of the last gather

    def _buf(self, r, nbytes):
        import torch
        if self.bufs[r] is None or self.bufs[r].numel() < nbytes:
            self.bufs[r] = torch.empty(int(nbytes * 1.1) + 4096, dtype=torch.uint8, device=self.dev)
        return self.bufs[r]

    def gather(self, out_ptr, runs_ptr, n_out_bytes, n_run_bytes, after_event=None, to_host=False):
        """Queues the gather on the gatherer's stream (after `after_event`, if given) -> total bytes that reach rank 0."""
        import torch
        import torch.distributed as dist
        with torch.cuda.stream(self.stream):
            if after_event is not None:
                self.stream.wait_event(after_event)
            out_d = torch.as_tensor(DevArr(out_ptr, max(1, n_out_bytes)), device=self.dev)[:n_out_bytes]
            runs_d = torch.as_tensor(DevArr(runs_ptr, max(1, n_run_bytes)), device=self.dev)[:n_run_bytes]
            sz = torch.tensor([n_out_bytes, n_run_bytes], dtype=torch.int64, device=self.dev)
            sizes = torch.empty(2 * self.world, dtype=torch.int64, device=self.dev)
            dist.all_gather_into_tensor(sizes, sz)
            sizes = sizes.cpu().numpy().reshape(self.world, 2)
            ops, parts = [], []
            if self.rank == 0:
                parts.append((out_d, runs_d))
                for r in range(1, self.world):
                    b = self._buf(r, int(sizes[r].sum()))
                    o, ru = b[:int(sizes[r, 0])], b[int(sizes[r, 0]):int(sizes[r].sum())]
                    if o.numel():
                        ops.append(dist.P2POp(dist.irecv, o, r))
                    if ru.numel():
                        ops.append(dist.P2POp(dist.irecv, ru, r))
                    parts.append((o, ru))
            else:
                if out_d.numel():
                    ops.append(dist.P2POp(dist.isend, out_d, 0))
                if runs_d.numel():
                    ops.append(dist.P2POp(dist.isend, runs_d, 0))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
            if self.rank == 0:
                self.last = parts
                if to_host:
                    tot = int(sizes.sum())
                    if self.host is None or self.host.numel() < tot:
                        self.host = torch.empty(int(tot * 1.1) + 4096, dtype=torch.uint8).pin_memory()
                    p = 0
                    for o, ru in parts:
                        for t in (o, ru):
                            self.host[p:p + t.numel()].copy_(t, non_blocking=True)
                            p += t.numel()
            return int(sizes.sum())
