#!/usr/bin/env python
"""bench.py -- anchored split-reads/s through the pattern-growth search (close end + far end) on N B200s.

  python bench.py --gpus N --steps K --warmup W                 the CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K --warmup W  the reference's own CPU implementation
                                                                  (oracle/_ref, built from the untouched sources)

One JSON line on stdout (rank 0).  A "step" is one pass of the search over one batch of synthetic anchored
reads (BASELINE.json configs[2]: 150 bp reads, 63 Mbp uniform-random reference, D/SI/INV/TD mix, Pindel defaults):
  value     reads/s with the batch already resident in HBM (K launches timed with CUDA events, max over ranks)
  e2e       reads/s through the C ABI with host buffers: H2D of the batch + search + D2H of the records per step
  roofline  algorithmic bytes per launch / measured launch time, against the measured HBM peak
  cpu_baseline  the reference's GetCloseEnd/SearchFarEnds on this box's host cores, on a bounded sample of the batch
Multi-GPU (torchrun, one rank per GPU): reads are sharded by 10 Mbp anchor window round-robin over the ranks
(weak scaling: the per-GPU batch is fixed), the packed reference is replicated, and every step ends with one NCCL
gather of the breakpoint records on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REF_BASES = 62_435_964          # test/SmallTest/hs_ref_chr20.fa.fai: "the 63 Mbp chr20-sized reference"
REF_SEED = 20260926
READ_LEN = 150
INSERT = 500
SHARD = 10_000_000              # anchor-window shard size for multi-GPU runs (BASELINE.json configs[3])


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step")
    ap.add_argument("--read-len", type=int, default=READ_LEN)
    ap.add_argument("--ref-bases", type=int, default=REF_BASES)
    ap.add_argument("--max-range-index", type=int, default=2, help="Pindel -x")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        # "under load": the upper half of the samples (idle samples before/after the region pull the median down)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "sm_mhz_min": min(sm) if sm else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def committed_traffic():
    """DRAM bytes per launch of the search kernel from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def workload_name(a, n_total):
    return ("synthetic %s anchored %dbp reads, %.0fMbp uniform-random reference (seed %d), D/SI/INV/TD 60/25/5/10, "
            "10%% decoys, 1%% errors, -x %d" % (("%dM" % (n_total // 1_000_000)) if n_total % 1_000_000 == 0 else str(n_total),
                                              a.read_len, a.ref_bases / 1e6, REF_SEED, a.max_range_index))


# ---------------------------------------------------------------------------------------------------------------
def make_batch_torch(ref_t, n, L, seed, windows):
    """Reads of this rank: the generator is run once per assigned window with a proportional share of n."""
    import torch
    from pindel_b200 import synth
    tot = sum(h - l for l, h in windows)
    parts, left = [], n
    for i, (l, h) in enumerate(windows):
        k = left if i == len(windows) - 1 else int(round(n * (h - l) / tot))
        k = min(k, left)
        left -= k
        if k > 0:
            parts.append(synth.make_reads_torch(ref_t, k, L, seed + 1000 * i, insert_size=INSERT, lo=l, hi=h))
    cat = [torch.cat([p[j] for p in parts]) for j in range(5)]
    return cat


class DevArr:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def algorithmic_compares(ref_np, seq_np, md_np, pos_np, isz_np, max_range_index, sample=5000):
    """CategorizePositions inner iterations per read (src/searcher.cpp:48-63), counted by the C port of the reference's
    algorithm (oracle/pg_oracle.c, pgo_char_compares) on the first `sample` reads with one thread."""
    import ctypes as C
    from oracle import pgoracle
    from pindel_b200 import synth
    lib = pgoracle.load()
    lib.pgo_char_compares.restype = C.c_ulonglong
    k = min(sample, seq_np.shape[0])
    L = seq_np.shape[1]
    P = pgoracle.default_params(max_range_index=max_range_index)
    lib.pgo_char_compares(1)
    pgoracle.search(P, [synth.pad_reference(ref_np)], seq_np[:k].reshape(-1), (np.arange(k + 1, dtype=np.uint64) * L).astype(np.uint32),
                    md_np[:k], pos_np[:k], isz_np[:k], threads=1)
    total = int(lib.pgo_char_compares(1))
    return {"char_compares_per_read": total / k, "unit": "Gcompare/s", "sample_reads": int(k)}


def cpu_baseline_reference(ref_np, seq_np, md_np, pos_np, isz_np, target_seconds, max_range_index, log):
    """Times the reference's own GetCloseEnd loop + SearchFarEnds (oracle/_ref/libpindel_ref.so) with all host
    threads on a bounded prefix of the batch.  Falls back to the C port when the reference library is absent."""
    from oracle import refhook, pgoracle
    from pindel_b200 import synth
    cores = os.cpu_count() or 1
    n_all = seq_np.shape[0]
    L = seq_np.shape[1]

    def blob(k):
        return seq_np[:k].reshape(-1), (np.arange(k + 1, dtype=np.uint64) * L).astype(np.uint32)

    if refhook.available():
        tmp = tempfile.mkdtemp(prefix="pgbench_")
        fa = os.path.join(tmp, "ref.fa")
        refhook.write_fasta(fa, [("chrS", ref_np)])
        h = refhook.RefHook(fa, ["-x", str(max_range_index), "-T", str(cores)], workdir=tmp)
        h.set_window(0, 0, ref_np.size)

        def run(k):
            b, o = blob(k)
            _, secs = h.search(b, o, md_np[:k], pos_np[:k], isz_np[:k], 0, True, cores, timing_only=True)
            return secs[0] + secs[1]
        kind = "reference"
    else:
        P = pgoracle.default_params(max_range_index=max_range_index)
        padded = synth.pad_reference(ref_np)

        def run(k):
            b, o = blob(k)
            _, secs = pgoracle.search(P, [padded], b, o, md_np[:k], pos_np[:k], isz_np[:k], threads=cores)
            return secs[0] + secs[1]
        kind = "port"
    k0 = min(n_all, 20000)
    t0 = run(k0)
    rate0 = k0 / max(t0, 1e-9)
    k = int(min(n_all, max(k0, rate0 * target_seconds)))
    t = run(k) if k > k0 else t0
    log("cpu_baseline(%s): %d reads in %.2fs on %d threads" % (kind, k, t, cores))
    return {"value": k / t, "unit": "reads/s", "cores": cores, "kind": kind,
            "sample": "first %d reads of the step's batch, GetCloseEnd loop + SearchFarEnds wall time, %d OpenMP threads" % (k, cores)}


# ---------------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def claim_stdout():
    """Only the result line may reach stdout: libraries print there too (NCCL's version banner, the reference's progress
    messages), so fd 1 is pointed at stderr for the whole run and the JSON line is written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(result):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(result) + "\n").encode())


def main():
    a = parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    def log(*x):
        print("[bench rank %d]" % rank, *x, file=sys.stderr, flush=True)

    if a.impl == "reference":
        return reference_arm(a, rank, world, log)

    import torch
    import torch.distributed as dist
    from pindel_b200 import api, runs as R, shard, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the search path has no CPU fallback (use --impl reference for the CPU reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n = a.reads
    L = a.read_len
    t0 = time.time()
    ref_np = synth.make_reference(a.ref_bases, REF_SEED)
    padded = synth.pad_reference(ref_np)
    ref_t = torch.from_numpy(ref_np).to(dev)
    seq_t, md_t, pos_t, isz_t, kind_t = make_batch_torch(ref_t, n, L, 777 + rank, shard.shard_windows(a.ref_bases, rank, world))
    del ref_t
    log("workload generated in %.1fs" % (time.time() - t0))

    pg = api.PatternGrowth(device=local_rank, max_range_index=a.max_range_index)
    # every copy and launch of the library goes on this (non-default) torch stream, so torch.cuda.Event times them
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    pg.set_stream(stream.cuda_stream)
    pg.set_reference([padded])

    # host (pinned) copies of the batch: what a caller of the C ABI holds
    seq_h = api.pinned_empty(n * L, np.uint8)
    off_h = api.pinned_empty(n + 1, np.uint32)
    md_h = api.pinned_empty(n, np.uint8)
    pos_h = api.pinned_empty(n, np.uint32)
    isz_h = api.pinned_empty(n, np.int16)
    seq_h[:] = seq_t.reshape(-1).cpu().numpy()
    off_h[:] = (np.arange(n + 1, dtype=np.uint64) * L).astype(np.uint32)
    md_h[:] = md_t.cpu().numpy()
    pos_h[:] = pos_t.cpu().numpy().astype(np.uint32)
    isz_h[:] = isz_t.cpu().numpy()
    del seq_t, md_t, pos_t, isz_t, kind_t
    torch.cuda.empty_cache()
    batch = dict(seq_blob=seq_h, seq_off=off_h, matched_d=md_h, matched_rel_pos=pos_h, insert_size=isz_h)
    h2d_bytes = seq_h.nbytes + off_h.nbytes + md_h.nbytes + pos_h.nbytes + isz_h.nbytes

    out_h = api.pinned_empty(n, R.OUT_DT)
    runs_h = api.pinned_empty(max(1024, 8 * n), R.RUN_DT)

    def gather_records():
        """The single collective of the path: breakpoint records (per-read headers + runs) to rank 0 over NCCL."""
        if world == 1:
            return
        optr, rptr, nruns = pg.device_results()
        out_d = torch.as_tensor(DevArr(optr, n * R.OUT_DT.itemsize), device=dev)
        runs_d = torch.as_tensor(DevArr(rptr, max(1, nruns * R.RUN_DT.itemsize)), device=dev)[:nruns * R.RUN_DT.itemsize]
        shard.gather_records(out_d, runs_d, dst=0)

    def step_resident():
        pg.launch(3)
        gather_records()

    def step_e2e():
        # the call a user makes: host buffers in, host buffers out (large batches are chunked and the copies overlap
        # the kernel inside the library); with several GPUs the records then go to rank 0 in one NCCL gather
        out_e, runs_e = pg.search(out=out_h, runs=runs_h, **batch)
        if world > 1:
            out_d = torch.from_numpy(out_e.view(np.uint8)).to(dev, non_blocking=True)
            runs_d = torch.from_numpy(runs_e.view(np.uint8)).to(dev, non_blocking=True)
            shard.gather_records(out_d, runs_d, dst=0)
        return out_e, runs_e

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    pg.upload(**batch)
    for _ in range(a.warmup):
        step_resident()
    pg.sync()
    window_bases, n_ups = pg.last_stats()
    launches_before = pg.kernel_launches

    # ---- timed: batch resident in HBM ----
    sampler = ClockSampler(local_rank)
    barrier()
    torch.cuda.profiler.start()        # ncu --profile-from-start off: only the timed regions are profiled
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(a.steps):
        step_resident()
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    launches = pg.kernel_launches - launches_before
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / a.steps
    value = world * n / (ms_step / 1e3)

    # ---- timed: end to end through the C ABI with host buffers ----
    for _ in range(min(2, a.warmup)):
        out_e, runs_e = step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        out_e, runs_e = step_e2e()
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    torch.cuda.profiler.stop()
    e2e_ms_step = float(t.item()) / a.steps
    e2e_value = world * n / (e2e_ms_step / 1e3)
    d2h_bytes = out_e.nbytes + runs_e.nbytes

    # ---- roofline of the search kernel (the only kernel in the timed region) ----
    read_bytes = n * ((L + 3) // 4 + 16)
    alg_bytes = read_bytes + window_bases // 4 + 16 * n_ups       # SURVEY.md section 8(d) / DESIGN.md section 5
    peak, peak_src = measured_hbm_peak()
    achieved = alg_bytes / (ms_step / 1e3) / 1e9
    traffic = committed_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (traffic["dram_bytes_per_read"] * n) if traffic and "dram_bytes_per_read" in traffic else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes),
                "bytes_per_read": alg_bytes / n, "window_bases_per_read": window_bases / n, "unique_points_per_read": n_ups / n,
                "note": "launch duration = CUDA-event time of the timed region / steps (includes the NCCL gather when gpus > 1); "
                        "the path is not HBM-bound at Pindel's default -x: the packed reference is L2-resident (see DESIGN.md)"}
    if traffic:
        roofline["traffic_source"] = traffic.get("source")
        if "compute" in traffic:
            roofline["compute"] = traffic["compute"]      # ncu issue/pipe utilisation of the same kernel (committed digest)

    result = {
        "metric": "anchored split-reads/sec through pattern-growth search", "value": value, "unit": "reads/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(a, world * n), "reads_per_gpu_per_step": n, "read_len": L,
                   "ref_bases": a.ref_bases, "insert_size": INSERT, "pindel_options": "defaults (-x %d -a 1 -m 3 -H 8 -u 0.02 -e 0.01 -E 0.95)" % a.max_range_index,
                   "parallelism": "reads sharded by 10 Mbp anchor window round-robin over %d GPU(s); reference replicated; one NCCL gather of the records per step" % world,
                   "l2": "inputs larger than L2: %.2f GB of read data stream through per step (L2 126 MB); the 23 MB packed reference is reused by design" % (h2d_bytes / 1e9)},
        "e2e": {"value": e2e_value, "unit": "reads/s", "ms_per_step": e2e_ms_step, "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": int(d2h_bytes)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
    }

    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            # op-roofline of SURVEY.md section 8(d)(ii): the character comparisons the reference's algorithm executes for these
            # reads (CategorizePositions inner iterations, counted by the oracle port on a small sample, one thread)
            ops = algorithmic_compares(ref_np, seq_h.reshape(n, L), md_h, pos_h, isz_h, a.max_range_index)
            ops["achieved_Gcompares_per_s"] = ops["char_compares_per_read"] * value / 1e9
            ops["note"] = ("algorithmic character comparisons per read (what the reference's level-bucket growth executes) x reads/s; the kernel "
                           "itself compares 32 bases per LOP3 pair on bit-planes and evaluates whole 32-base stretches per seed, so it executes "
                           "more raw comparisons than this count -- this figure is the useful work, comparable across implementations")
            roofline["ops"] = ops
        except Exception as e:
            roofline["ops"] = {"unavailable": repr(e)}
        try:
            result["cpu_baseline"] = cpu_baseline_reference(ref_np, seq_h.reshape(n, L), md_h, pos_h, isz_h, a.cpu_seconds,
                                                            a.max_range_index, log)
        except Exception as e:  # the baseline is reported, never required for the GPU numbers
            result["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(e)}
    if rank == 0:
        emit(result)
    pg.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def reference_arm(a, rank, world, log):
    """The reference's own CPU implementation of the path (GetCloseEnd + SearchFarEnds from oracle/_ref, built from the
    untouched sources), all host threads, same metric/config; each step is a bounded sample of the workload."""
    if rank != 0:
        return
    from pindel_b200 import synth
    from oracle import refhook, pgoracle
    L = a.read_len
    cores = os.cpu_count() or 1
    ref_np = synth.make_reference(a.ref_bases, REF_SEED)
    # bounded sample per step, sized so that K + W steps end within a few minutes on ~10 k reads/s hosts
    n_step = max(2000, min(a.reads, int(120_000 / max(1, a.steps + a.warmup))))
    batch = synth.make_reads(ref_np, n_step, L, seed=777, insert_size=INSERT)
    blob, off = batch.blob()
    if refhook.available():
        tmp = tempfile.mkdtemp(prefix="pgbench_")
        fa = os.path.join(tmp, "ref.fa")
        refhook.write_fasta(fa, [("chrS", ref_np)])
        h = refhook.RefHook(fa, ["-x", str(a.max_range_index), "-T", str(cores)], workdir=tmp)
        h.set_window(0, 0, ref_np.size)

        def run():
            _, secs = h.search(blob, off, batch.matched_d, batch.matched_rel_pos, batch.insert_size, 0, True, cores, timing_only=True)
            return secs[0] + secs[1]
        kind = "reference"
    else:
        P = pgoracle.default_params(max_range_index=a.max_range_index)
        padded = synth.pad_reference(ref_np)

        def run():
            _, secs = pgoracle.search(P, [padded], blob, off, batch.matched_d, batch.matched_rel_pos, batch.insert_size, threads=cores)
            return secs[0] + secs[1]
        kind = "port"
    for _ in range(a.warmup):
        run()
    tot = 0.0
    for _ in range(a.steps):
        tot += run()
    ms_step = tot / a.steps * 1e3
    value = n_step / (ms_step / 1e3)
    sample = "%d reads per step (bounded sample of the %d-read workload), GetCloseEnd loop + SearchFarEnds wall time, %d OpenMP threads" % (n_step, a.reads * max(1, world), cores)
    result = {
        "impl": "reference",
        "metric": "anchored split-reads/sec through pattern-growth search", "value": value, "unit": "reads/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(a, max(1, world) * a.reads), "reads_per_step_sampled": n_step, "read_len": L,
                   "ref_bases": a.ref_bases, "insert_size": INSERT,
                   "pindel_options": "defaults (-x %d -a 1 -m 3 -H 8 -u 0.02 -e 0.01 -E 0.95)" % a.max_range_index,
                   "parallelism": "OpenMP, %d host threads (rank 0 only)" % cores},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(result)


if __name__ == "__main__":
    main()
