#!/usr/bin/env python
"""bench.py -- anchored split-reads/s through the pattern-growth search (close end + far end) on N B200s.

  python bench.py --gpus N --steps K --warmup W                 the CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K --warmup W  the reference's own CPU implementation
                                                                  (oracle/_ref, built from the untouched sources)

One JSON line on stdout (rank 0).  A "step" is one pass of the search over one batch of synthetic anchored reads.  The
headline workload is BASELINE.json configs[2] (10 M x 150 bp reads, 62 Mbp uniform-random reference, D/SI/INV/TD mix,
Pindel defaults):
  value     reads/s with the batch already resident in HBM (K launches timed with CUDA events, max over ranks; with
            several GPUs every step ends with the NCCL gather of the breakpoint records on rank 0, which travels while the
            next step's kernel runs)
  e2e       reads/s through the C ABI with host buffers: H2D of the batch + search + D2H of the records per step (with
            several GPUs every rank's result buffers are a shared-memory segment that rank 0 maps)
  roofline  algorithmic bytes per launch / measured launch time, against the measured HBM peak
  parity    after the timed region: a sample of the last step's records (from every rank) against the CPU restatement
  cpu_baseline  the reference's GetCloseEnd/SearchFarEnds on this box's host cores, on a bounded sample of the batch,
            and the whole reference program (`pindel_ref -T nproc`) with its own phase timers on the same reads
  secondary the other shapes of BASELINE.json, each measured the same way in the same run: configs[2] at -x 5 (the
            large-window regime), configs[1] (1 M x 100 bp, 10 Mbp, deletions only), configs[4]'s shape (3 Gbp / 24 contigs,
            reads sharded by chromosome) and, with several GPUs, configs[3] (50 M reads strong-scaled over 10 Mbp shards)
Multi-GPU (torchrun, one rank per GPU): reads are sharded by anchor window (or chromosome) over the ranks, the packed
reference is replicated, and the only collective is the gather of the records.

oracle/ is test infrastructure: this file uses it only as the checker (`parity`, after the timed region) and as the CPU
baseline beside the GPU numbers, never as the thing measured.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REF_BASES = 62_435_964          # test/SmallTest/hs_ref_chr20.fa.fai: "the 63 Mbp chr20-sized reference"
REF_SEED = 20260926
READ_LEN = 150
INSERT = 500
SHARD = 10_000_000              # anchor-window shard size for multi-GPU runs (BASELINE.json configs[3])
# 24 contigs of a 3.1 Gbp genome (human-like proportions, in Mbp), BASELINE.json configs[4]
CONTIGS_MBP = [248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80, 58, 64, 46, 50, 156, 57]
WG_SEED = 20260927
METRIC = "anchored split-reads/sec through pattern-growth search"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step (headline workload)")
    ap.add_argument("--read-len", type=int, default=READ_LEN)
    ap.add_argument("--ref-bases", type=int, default=REF_BASES)
    ap.add_argument("--max-range-index", type=int, default=2, help="Pindel -x of the headline workload")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of each cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--secondary", default="all", help="all | none | comma list of x5,c2,wg,strong")
    ap.add_argument("--wg-reads", type=int, default=20_000_000, help="reads per GPU of the 3 Gbp workload")
    ap.add_argument("--parity-sample", type=int, default=20_000)
    ap.add_argument("--cpu-worker", default=None, help=argparse.SUPPRESS)
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


def committed_traffic(key):
    """DRAM bytes per read of the search kernel from the committed ncu --set full captures (profiles/traffic.json), per
    workload key; a static figure from a capture, not a measurement of this run."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get(key)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------------
# Workloads.  Every workload is described by a spec (plain dict, printed as `config`) and generated by make_workload on
# the rank's GPU with torch ops (the synthesis is not part of any timed region); the reference arm builds the very same
# arrays with the same call.
# ---------------------------------------------------------------------------------------------------------------
def spec_c3(a, world, x=None, reads=None, scaling="weak"):
    x = a.max_range_index if x is None else x
    reads = a.reads if reads is None else reads
    tot = world * reads
    return {"key": "c3_x%d" % x,
            "workload": "synthetic %s anchored %dbp reads, %.0fMbp uniform-random reference (seed %d), D/SI/INV/TD 60/25/5/10, 10%% decoys, "
                        "1%% errors, -x %d" % (("%dM" % (tot // 1_000_000)) if tot % 1_000_000 == 0 else str(tot), a.read_len,
                                               a.ref_bases / 1e6, REF_SEED, x),
            "kind": "single", "reads_per_gpu_per_step": reads, "read_len": a.read_len, "ref_bases": a.ref_bases, "ref_seed": REF_SEED,
            "insert_size": INSERT, "max_range_index": x, "mix": [0.60, 0.25, 0.05, 0.10], "scaling": scaling,
            "pindel_options": "defaults (-x %d -a 1 -m 3 -H 8 -u 0.02 -e 0.01 -E 0.95)" % x,
            "sharding": "reads sharded by 10 Mbp anchor window round-robin over the GPUs; reference replicated"}


def spec_c2(a):
    return {"key": "c2", "workload": "synthetic 1M anchored 100bp reads, 10Mbp uniform-random reference (seed 20260925), deletions only, 10% decoys, 1% errors, -x 2",
            "kind": "single", "reads_per_gpu_per_step": 1_000_000, "read_len": 100, "ref_bases": 10_000_000, "ref_seed": 20260925,
            "insert_size": INSERT, "max_range_index": 2, "mix": [1.0, 0.0, 0.0, 0.0], "scaling": "weak",
            "pindel_options": "defaults (-x 2 -a 1 -m 3 -H 8 -u 0.02 -e 0.01 -E 0.95)",
            "sharding": "reads sharded by 10 Mbp anchor window round-robin over the GPUs; reference replicated"}


def spec_wg(a, world):
    tot = world * a.wg_reads
    return {"key": "wg", "workload": "synthetic %dM anchored 150bp reads, %.2f Gbp uniform-random reference in 24 contigs (seed %d), D/SI/INV/TD 60/25/5/10, "
                                     "10%% decoys, 1%% errors, -x 2" % (tot // 1_000_000, sum(CONTIGS_MBP) / 1e3, WG_SEED),
            "kind": "genome", "reads_per_gpu_per_step": a.wg_reads, "read_len": 150, "contigs_mbp": CONTIGS_MBP, "ref_bases": sum(CONTIGS_MBP) * 1_000_000,
            "ref_seed": WG_SEED, "insert_size": INSERT, "max_range_index": 2, "mix": [0.60, 0.25, 0.05, 0.10], "scaling": "weak",
            "pindel_options": "defaults (-x 2 -a 1 -m 3 -H 8 -u 0.02 -e 0.01 -E 0.95)",
            "sharding": "reads sharded by chromosome over the GPUs (longest first to the least loaded rank); 1.1 GB packed reference replicated"}


def config_of(spec, world):
    """The `config` object of the JSON line: identical in both arms (what differs between them sits outside it)."""
    n, L = spec["reads_per_gpu_per_step"], spec["read_len"]
    cfg = {k: v for k, v in spec.items() if k not in ("key", "kind", "mix", "ref_seed", "contigs_mbp", "sharding")}
    cfg["parallelism"] = "%s; one NCCL gather of the records per step, overlapped with the next step's kernel (%d GPU(s))" % (spec["sharding"], world)
    packed_mb = (spec["ref_bases"] + 200_000 * (len(spec.get("contigs_mbp", [0])))) * 3 / 8 / 1e6
    cfg["l2"] = "inputs larger than L2: %.2f GB of read data stream through per GPU per step (L2 126 MB); packed reference %.0f MB (%s)" % (
        n * (L + 11) / 1e9, packed_mb, "L2-resident, reused by design" if packed_mb < 100 else "larger than L2")
    return cfg


def contig_owner(world):
    """Chromosome -> rank: longest first to the least loaded rank (per-chromosome shards, BASELINE.json configs[4])."""
    load = [0] * world
    owner = [0] * len(CONTIGS_MBP)
    for c in sorted(range(len(CONTIGS_MBP)), key=lambda i: -CONTIGS_MBP[i]):
        r = min(range(world), key=lambda j: load[j])
        owner[c] = r
        load[r] += CONTIGS_MBP[c]
    return owner


def make_workload(spec, rank, world, dev):
    """-> dict(padded=[uint8 arrays], seq [n,L] uint8, md, pos uint32, isz int16, chr_id int32 | None); host numpy arrays.
    Generated with torch on `dev` (a CUDA device, or the CPU when there is none)."""
    import torch
    from pindel_b200 import shard, synth
    n, L = spec["reads_per_gpu_per_step"], spec["read_len"]
    mix = tuple(spec["mix"])
    if spec["kind"] == "single":
        ref_np = synth.make_reference(spec["ref_bases"], spec["ref_seed"])
        ref_t = torch.from_numpy(ref_np).to(dev)
        wins = shard.shard_windows(spec["ref_bases"], rank, world)
        tot = sum(h - l for l, h in wins)
        parts, left = [], n
        for i, (l, h) in enumerate(wins):
            k = left if i == len(wins) - 1 else min(left, int(round(n * (h - l) / tot)))
            left -= k
            if k > 0:
                parts.append(synth.make_reads_torch(ref_t, k, L, 777 + rank + 1000 * i, insert_size=spec["insert_size"], mix=mix, lo=l, hi=h))
        cat = [torch.cat([p[j] for p in parts]) for j in range(4)]
        return dict(padded=[synth.pad_reference(ref_np)], seq=cat[0].cpu().numpy(), md=cat[1].cpu().numpy(),
                    pos=cat[2].cpu().numpy().astype(np.uint32), isz=cat[3].cpu().numpy(), chr_id=None)
    # whole genome: every rank holds all contigs, generates reads on its own
    owner = contig_owner(world)
    mine = [c for c in range(len(CONTIGS_MBP)) if owner[c] == rank] or [rank % len(CONTIGS_MBP)]
    tot = sum(CONTIGS_MBP[c] for c in mine)
    gen = torch.Generator(device=dev)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    padded, parts, chr_parts, left = [], [], [], n
    for c, mbp in enumerate(CONTIGS_MBP):
        gen.manual_seed(spec["ref_seed"] * 100 + c)
        ref_t = acgt[torch.randint(0, 4, (mbp * 1_000_000,), generator=gen, device=dev, dtype=torch.uint8).long()]
        pad = np.full(ref_t.numel() + 2 * synth.SPACER, ord("N"), dtype=np.uint8)
        pad[synth.SPACER:synth.SPACER + ref_t.numel()] = ref_t.cpu().numpy()
        padded.append(pad)
        if c in mine:
            k = left if c == mine[-1] else min(left, int(round(n * mbp / tot)))
            left -= k
            if k > 0:
                parts.append(synth.make_reads_torch(ref_t, k, L, 4242 + c, insert_size=spec["insert_size"], mix=mix))
                chr_parts.append(np.full(k, c, dtype=np.int32))
        del ref_t
    cat = [torch.cat([p[j] for p in parts]) for j in range(4)]
    return dict(padded=padded, seq=cat[0].cpu().numpy(), md=cat[1].cpu().numpy(), pos=cat[2].cpu().numpy().astype(np.uint32),
                isz=cat[3].cpu().numpy(), chr_id=np.concatenate(chr_parts))


# ---------------------------------------------------------------------------------------------------------------
# CPU legs.  The reference keeps process-global settings, so every configuration runs in its own process: this file again,
# with --cpu-worker <dir> (inputs as .npy files, the answer as one JSON line on stdout).
# ---------------------------------------------------------------------------------------------------------------
def cpu_worker(wdir):
    from oracle import refhook, pgoracle
    from pindel_b200 import synth
    cfg = json.load(open(os.path.join(wdir, "cfg.json")))
    seq = np.load(os.path.join(wdir, "seq.npy"), mmap_mode="r")
    md, pos, isz = (np.load(os.path.join(wdir, k + ".npy")) for k in ("md", "pos", "isz"))
    ref = np.load(os.path.join(wdir, "ref.npy"))          # unpadded ASCII of the one chromosome the sample lies on
    n_all, L = seq.shape
    cores = os.cpu_count() or 1
    x = cfg["max_range_index"]

    def blob(k):
        return np.ascontiguousarray(seq[:k]).reshape(-1), (np.arange(k + 1, dtype=np.uint64) * L).astype(np.uint32)

    if refhook.available():
        fa = os.path.join(wdir, "ref.fa")
        refhook.write_fasta(fa, [("chrS", ref)])
        h = refhook.RefHook(fa, ["-x", str(x), "-T", str(cores)], workdir=wdir)
        h.set_window(0, 0, ref.size)

        def run(k):
            b, o = blob(k)
            _, secs = h.search(b, o, md[:k], pos[:k], isz[:k], 0, True, cores, timing_only=True)
            return secs[0] + secs[1]
        kind = "reference"
    else:
        P = pgoracle.default_params(max_range_index=x)
        padded = synth.pad_reference(ref)

        def run(k):
            b, o = blob(k)
            _, secs = pgoracle.search(P, [padded], b, o, md[:k], pos[:k], isz[:k], threads=cores)
            return secs[0] + secs[1]
        kind = "port"
    k0 = min(n_all, 20000)
    t0 = run(k0)
    k = int(min(n_all, max(k0, k0 / max(t0, 1e-9) * cfg["seconds"])))
    out = {"kind": kind, "cores": cores, "probe_reads": k0, "probe_seconds": t0}
    if cfg.get("steps"):                      # the reference arm: W + K steps of k reads each
        for _ in range(cfg["warmup"]):
            run(k)
        ts = [run(k) for _ in range(cfg["steps"])]
        out.update(reads=k, seconds=float(np.mean(ts)), step_seconds=ts)
    else:
        t = run(k) if k > k0 else t0
        out.update(reads=k, seconds=t)
    # SURVEY.md section 8(d) leg (1): the whole reference program on the same reads as a Pindel-text file, its own phase timers
    binary = os.path.join(ROOT, "oracle", "_ref", "pindel_ref")
    if cfg.get("whole_program") and os.path.exists(binary):
        try:
            kk = int(min(n_all, max(k0, k0 / max(t0, 1e-9) * cfg["whole_program"])))
            fa = os.path.join(wdir, "wp.fa")
            refhook.write_fasta(fa, [("chrS", ref)])
            rows = 0 if ref.size % 60 == 0 else 1
            open(fa + ".fai", "w").write("chrS\t%d\t6\t60\t61\n" % ref.size)
            batch = synth.ReadBatch(np.ascontiguousarray(seq[:kk]), md[:kk], pos[:kk], isz[:kk], np.zeros(kk, np.uint8))
            open(os.path.join(wdir, "reads.txt"), "w").write(batch.to_pindel_text("chrS"))
            t = time.time()
            p = subprocess.run([binary, "-f", fa, "-p", os.path.join(wdir, "reads.txt"), "-o", os.path.join(wdir, "wp.out"), "-x", str(x),
                                "-T", str(cores)], capture_output=True, text=True, timeout=900)
            wall = time.time() - t
            phases = {m.group(1).strip(): int(m.group(2)) for m in re.finditer(r"^(.*?) (\d+) seconds\.$", p.stdout, re.M)}
            out["whole_program"] = {"command": "pindel_ref -f ref.fa -p reads.txt -x %d -T %d" % (x, cores), "reads": kk, "wall_seconds": wall,
                                    "reads_per_s": kk / wall, "phases_seconds": phases, "rc": p.returncode,
                                    "note": "the reference's own Timer has 1 s resolution (src/pindel.cpp:1177-1260)"}
        except Exception as e:
            out["whole_program"] = {"unavailable": repr(e)}
    sys.stdout.write(json.dumps(out) + "\n")


def run_cpu_leg(w, spec, seconds, log, steps=0, warmup=0, whole_program=0.0, max_reads=2_000_000):
    """Writes the (bounded) sample of workload `w` for the worker and returns its answer.  Samples come from ONE chromosome
    (the first reads' one), because the worker loads a single-contig reference."""
    n = min(w["seq"].shape[0], max_reads)
    sel = slice(0, n)
    if w["chr_id"] is not None:
        c = int(w["chr_id"][0])
        idx = np.nonzero(w["chr_id"] == c)[0][:max_reads]
        sel = idx
        ref = w["padded"][c][100000:-100000]
    else:
        ref = w["padded"][0][100000:-100000]
    wdir = tempfile.mkdtemp(prefix="pgbench_cpu_")
    try:
        np.save(os.path.join(wdir, "seq.npy"), w["seq"][sel])
        for k in ("md", "pos", "isz"):
            np.save(os.path.join(wdir, k + ".npy"), w[k][sel])
        np.save(os.path.join(wdir, "ref.npy"), ref)
        json.dump({"max_range_index": spec["max_range_index"], "seconds": seconds, "steps": steps, "warmup": warmup,
                   "whole_program": whole_program}, open(os.path.join(wdir, "cfg.json"), "w"))
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-worker", wdir], capture_output=True, text=True, timeout=3000)
        if p.returncode != 0:
            raise RuntimeError("cpu worker failed: " + p.stderr[-800:])
        ans = json.loads(p.stdout.strip().split("\n")[-1])
    finally:
        shutil.rmtree(wdir, ignore_errors=True)
    log("cpu leg (%s, %s): %d reads in %.2fs on %d threads" % (spec["key"], ans["kind"], ans["reads"], ans["seconds"], ans["cores"]))
    return ans


def cpu_baseline_entry(ans):
    cb = {"value": ans["reads"] / ans["seconds"], "unit": "reads/s", "cores": ans["cores"], "kind": ans["kind"],
          "sample": "first %d reads of the step's batch, GetCloseEnd loop + SearchFarEnds wall time, %d OpenMP threads" % (ans["reads"], ans["cores"])}
    if "whole_program" in ans:
        cb["phases"] = ans["whole_program"]
    return cb


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


def algorithmic_compares(w, spec, sample=5000):
    if spec["max_range_index"] >= 4:
        sample = 1000            # one thread of the level-bucket growth manages a few hundred reads/s at -x 5
    """CategorizePositions inner iterations per read (src/searcher.cpp:48-63), counted by the C port of the reference's
    algorithm (oracle/pg_oracle.c, pgo_char_compares) on the first `sample` reads with one thread."""
    import ctypes as C
    from oracle import pgoracle
    lib = pgoracle.load()
    lib.pgo_char_compares.restype = C.c_ulonglong
    k = min(sample, w["seq"].shape[0])
    L = w["seq"].shape[1]
    P = pgoracle.default_params(max_range_index=spec["max_range_index"])
    lib.pgo_char_compares(1)
    pgoracle.search(P, w["padded"], np.ascontiguousarray(w["seq"][:k]).reshape(-1), (np.arange(k + 1, dtype=np.uint64) * L).astype(np.uint32),
                    w["md"][:k], w["pos"][:k], w["isz"][:k], chr_id=None if w["chr_id"] is None else w["chr_id"][:k], threads=1)
    total = int(lib.pgo_char_compares(1))
    return {"char_compares_per_read": total / k, "unit": "Gcompare/s", "sample_reads": int(k)}


# ---------------------------------------------------------------------------------------------------------------
class SharedRecords:
    """Result buffers of one rank (pg_read_out[n] + pg_run[cap]) in a /dev/shm file, page-locked for the GPU's copies; rank 0 maps
    every rank's file, which makes the end-to-end "gather" a matter of each rank's own D2H."""

    def __init__(self, rank, n, run_cap, torch):
        from pindel_b200 import runs as R
        self.torch, self.n, self.run_cap = torch, n, run_cap
        self.nbytes = n * R.OUT_DT.itemsize + run_cap * R.RUN_DT.itemsize
        # /dev/shm when it has room for every rank's segment, else a file under the temporary directory (page cache)
        base = "/dev/shm"
        try:
            st = os.statvfs(base)
            if st.f_blocks * st.f_frsize < 12 * self.nbytes:      # capacity, not free space: every rank must decide alike
                base = tempfile.gettempdir()
        except OSError:
            base = tempfile.gettempdir()
        self.tag = os.path.join(base, "pgbench_%s_" % os.environ.get("MASTER_PORT", "0"))
        self.path = "%s%d" % (self.tag, rank)
        self.mm = np.memmap(self.path, dtype=np.uint8, mode="w+", shape=(self.nbytes,))
        self.mm[:] = 0                                    # touch every page before registering
        self.registered = torch.cuda.cudart().cudaHostRegister(self.mm.ctypes.data, self.nbytes, 0)
        self.out = self.mm[:n * R.OUT_DT.itemsize].view(R.OUT_DT)
        self.runs = self.mm[n * R.OUT_DT.itemsize:].view(R.RUN_DT)

    def check_on_rank0(self, world, gat, R):
        """Rank 0 maps the other ranks' segments and compares the first headers with what the NCCL gather of the resident path
        delivered (same reads, same kernel: lengths, flags and run counts must agree).  -> ranks seen"""
        import torch.distributed as dist
        dist.barrier()
        if dist.get_rank() != 0:
            return None
        seen = 0
        for r in range(world):
            mm = np.memmap("%s%d" % (self.tag, r), dtype=np.uint8, mode="r")
            k = min(1000, self.n)
            mine = mm[:k * R.OUT_DT.itemsize].view(R.OUT_DT)
            ref = gat.last[r][0][:k * R.OUT_DT.itemsize].cpu().numpy().view(R.OUT_DT)
            if all(np.array_equal(mine[f], ref[f]) for f in ("n_close_runs", "n_far_runs", "len_after", "flags")):
                seen += 1
            del mm
        return seen

    def close(self):
        try:
            self.torch.cuda.cudart().cudaHostUnregister(self.mm.ctypes.data)
        except Exception:
            pass
        del self.out, self.runs
        try:
            self.mm._mmap.close()
        except Exception:
            pass
        try:
            os.unlink(self.path)
        except OSError:
            pass


def run_workload(a, spec, w, rank, local_rank, world, dev, steps, warmup, log, want_cpu):
    """Times one workload on this rank's GPU (collectively with the other ranks) -> result dict (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from pindel_b200 import api, runs as R, shard

    n, L = w["seq"].shape
    pg = api.PatternGrowth(device=local_rank, max_range_index=spec["max_range_index"])
    # every copy and launch of the library goes on this (non-default) torch stream, so torch.cuda.Event times them
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    pg.set_stream(stream.cuda_stream)
    pg.set_reference(w["padded"])
    # host (pinned) copies of the batch: what a caller of the C ABI holds
    seq_h = api.pinned_empty(n * L, np.uint8)
    off_h = api.pinned_empty(n + 1, np.uint32)
    md_h = api.pinned_empty(n, np.uint8)
    pos_h = api.pinned_empty(n, np.uint32)
    isz_h = api.pinned_empty(n, np.int16)
    seq_h[:] = w["seq"].reshape(-1)
    off_h[:] = (np.arange(n + 1, dtype=np.uint64) * L).astype(np.uint32)
    md_h[:] = w["md"]; pos_h[:] = w["pos"]; isz_h[:] = w["isz"]
    batch = dict(seq_blob=seq_h, seq_off=off_h, matched_d=md_h, matched_rel_pos=pos_h, insert_size=isz_h)
    h2d_bytes = seq_h.nbytes + off_h.nbytes + md_h.nbytes + pos_h.nbytes + isz_h.nbytes
    if w["chr_id"] is not None:
        chr_h = api.pinned_empty(n, np.int32)
        chr_h[:] = w["chr_id"]
        batch["chr_id"] = chr_h
        h2d_bytes += chr_h.nbytes
    shm = None
    if world == 1:
        out_h = api.pinned_empty(n, R.OUT_DT)
        runs_h = api.pinned_empty(max(1024, 8 * n), R.RUN_DT)
    else:
        # several GPUs, end to end: every rank's records land in a host shared-memory segment that rank 0 maps (page-locked with
        # cudaHostRegister), so the records of all ranks are in rank 0's address space after each rank's own D2H -- N PCIe links
        # in parallel instead of one
        shm = SharedRecords(rank, n, 4 * n + 4096, torch)
        out_h, runs_h = shm.out, shm.runs
    gat = shard.RecordGatherer(dev, rank, world) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def resident_steps(k):
        """k launches over the resident batch; with several GPUs the records of launch i travel to rank 0 while launch
        i+1 runs (the library alternates between two result slots)."""
        evs, gdone = [], []
        for i in range(k):
            if gat and i >= 3:
                stream.wait_event(gdone[i - 3])     # launch i reuses the result slot of launch i-3: its gather comes first
            pg.launch(3)
            if gat:
                ev = torch.cuda.Event()
                ev.record(stream)
                evs.append(ev)
                if i > 0:
                    o, r, nr, nn = pg.device_results_previous()
                    gat.gather(o, r, nn * R.OUT_DT.itemsize, nr * R.RUN_DT.itemsize, after_event=evs[i - 1])
                    g = torch.cuda.Event()
                    g.record(gat.stream)
                    gdone.append(g)
        if gat:
            o, r, nr = pg.device_results()
            gat.gather(o, r, n * R.OUT_DT.itemsize, nr * R.RUN_DT.itemsize, after_event=evs[-1])
            stream.wait_stream(gat.stream)

    def e2e_steps(k):
        """The call a user makes, k times: host buffers in, host buffers out (pg_search; large batches are chunked so the copies
        overlap the kernel).  With several GPUs every rank does that for its shard over its own PCIe link, and its result
        buffers are a shared-memory segment rank 0 maps: after the closing barrier all records are in rank 0's address space.
        (The first version of this round gathered the records on the device -- NCCL, as the resident path does -- and copied them
        to the host on rank 0: 3.9 GB per step through ONE PCIe link at N = 8, 334 M reads/s instead of ~600,
        profiles/r2q_bench_n8_gather_then_d2h.json.)"""
        res = None
        for _ in range(k):
            res = pg.search(out=out_h, runs=runs_h, **batch)
        moved = res[0].nbytes + res[1].nbytes
        if world > 1:
            t = torch.tensor([moved], device=dev, dtype=torch.int64)
            dist.all_reduce(t)
            moved = int(t.item())
        return moved

    # ---- warm-up ----
    pg.upload(**batch)
    resident_steps(max(1, warmup))
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
    resident_steps(steps)
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_step = max_over_ranks(ev0.elapsed_time(ev1)) / steps
    launches = pg.kernel_launches - launches_before
    value = world * n / (ms_step / 1e3)

    # ---- parity: a sample of the last step's records, from every rank, against the CPU restatement ----
    parity = parity_check(a, spec, w, pg, gat, rank, world, n, L, log, out_h, runs_h, dev)

    # ---- timed: end to end through the C ABI with host buffers ----
    e2e_steps(min(2, max(1, warmup)))
    barrier()
    launches_e2e0 = pg.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    d2h_bytes = e2e_steps(steps)
    e1.record(stream)
    barrier()
    torch.cuda.profiler.stop()
    e2e_ms_step = max_over_ranks(e0.elapsed_time(e1)) / steps
    e2e_value = world * n / (e2e_ms_step / 1e3)
    e2e_seen = shm.check_on_rank0(world, gat, R) if shm else None
    launches += pg.kernel_launches - launches_e2e0

    # ---- roofline of the search kernel (the only kernel in the timed region) ----
    read_bytes = n * ((L + 3) // 4 + 16)
    alg_bytes = read_bytes + window_bases // 4 + 16 * n_ups       # SURVEY.md section 8(d) / DESIGN.md section 5
    peak, peak_src = measured_hbm_peak()
    achieved = alg_bytes / (ms_step / 1e3) / 1e9
    traffic = committed_traffic(spec["key"])
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (traffic["dram_bytes_per_read"] * n) if traffic and "dram_bytes_per_read" in traffic else None,
                "traffic_kind": "static: committed ncu capture scaled per read, not a measurement of this run" if traffic else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes),
                "bytes_per_read": alg_bytes / n, "window_bases_per_read": window_bases / n, "unique_points_per_read": n_ups / n,
                "note": "launch duration = CUDA-event time of the timed region / steps; the path is not HBM-bound while the packed reference fits "
                        "the 126 MB L2 (see DESIGN.md section 5): the binding resource is instruction issue"}
    if traffic:
        roofline["traffic_source"] = traffic.get("source")
        if "compute" in traffic:
            roofline["compute"] = traffic["compute"]      # ncu issue/pipe utilisation of the same kernel (committed digest)
    cfg = config_of(spec, world)
    result = {"metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
              "higher_is_better": True, "scaling": spec["scaling"], "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": cfg,
              "e2e": {"value": e2e_value, "unit": "reads/s", "ms_per_step": e2e_ms_step, "h2d_bytes_per_step": int(h2d_bytes * world),
                      "d2h_bytes_per_step": int(d2h_bytes), **({"records_visible_on_rank0": e2e_seen} if e2e_seen is not None else {})},
              "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "parity": parity}
    if rank == 0 and want_cpu:
        try:
            ops = algorithmic_compares(w, spec)
            ops["achieved_Gcompares_per_s"] = ops["char_compares_per_read"] * value / 1e9
            ops["note"] = ("algorithmic character comparisons per read (the inner statement of the reference's level-bucket growth, counted by the "
                           "oracle port on a sample) x reads/s: the useful work, comparable across implementations")
            roofline["ops"] = ops
        except Exception as e:
            roofline["ops"] = {"unavailable": repr(e)}
        try:
            ans = run_cpu_leg(w, spec, a.cpu_seconds, log, whole_program=(8.0 if spec["key"].startswith("c3_x%d" % a.max_range_index) else 0.0))
            result["cpu_baseline"] = cpu_baseline_entry(ans)
        except Exception as e:  # the baseline is reported, never required for the GPU numbers
            result["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(e)}
    pg.close()
    del seq_h, off_h, md_h, pos_h, isz_h, out_h, runs_h
    if shm:
        if world > 1:
            dist.barrier()
        shm.close()
    torch.cuda.empty_cache()
    return result if rank == 0 else None


def parity_check(a, spec, w, pg, gat, rank, world, n, L, log, out_h, runs_h, dev):
    """Rank 0 compares the first k reads of every rank's last launch with the oracle port (oracle/pg_oracle.c, pinned to the
    reference): the records are the ones the gather delivered (merged with shard.merge_gathered), the inputs of the other
    ranks' samples travel as objects.  -> {"checked", "mismatches", "ranks"}"""
    import torch.distributed as dist
    from pindel_b200 import runs as R, shard
    k = min(n, max(1, (a.parity_sample + world - 1) // world))
    try:
        mine = dict(seq=np.ascontiguousarray(w["seq"][:k]), md=w["md"][:k].copy(), pos=w["pos"][:k].copy(), isz=w["isz"][:k].copy(),
                    chr_id=None if w["chr_id"] is None else w["chr_id"][:k].copy())
        if world > 1:
            samples = [None] * world if rank == 0 else None
            dist.gather_object(mine, samples, dst=0)
        else:
            samples = [mine]
        if rank != 0:
            return None
        import torch

        def sample_of(out_k, take_runs):
            """The first k headers with their runs gathered into one compact array (run_off rewritten): the runs of a read
            sit wherever the kernel's allocator put them."""
            cnt = out_k["n_close_runs"].astype(np.int64) + out_k["n_far_runs"].astype(np.int64)
            new_off = np.cumsum(cnt) - cnt
            idx = np.repeat(out_k["run_off"].astype(np.int64) - new_off, cnt) + np.arange(int(cnt.sum()), dtype=np.int64)
            o = out_k.copy()
            o["run_off"] = new_off.astype(np.uint32)
            return o, take_runs(idx)

        if gat:
            parts = []
            for o, ru in gat.last:
                out_k = o[:k * R.OUT_DT.itemsize].cpu().numpy().view(R.OUT_DT)
                rows = ru.view(-1, R.RUN_DT.itemsize)
                parts.append(sample_of(out_k, lambda idx: rows[torch.from_numpy(idx).to(dev)].cpu().numpy().reshape(-1).view(R.RUN_DT)))
        else:
            out, runs = pg.download(out=out_h, runs=runs_h)
            parts = [sample_of(out[:k], lambda idx: runs[idx])]
        owners = np.repeat(np.arange(world, dtype=np.int32), k)
        mo, mr = shard.merge_gathered(parts, owners)
        seq = np.concatenate([s["seq"] for s in samples])
        blob, off = seq.reshape(-1), (np.arange(seq.shape[0] + 1, dtype=np.uint64) * L).astype(np.uint32)
        got = R.per_read(mo, mr, blob, off)
        from oracle import pgoracle, refhook
        P = pgoracle.default_params(max_range_index=spec["max_range_index"])
        chr_id = None if samples[0]["chr_id"] is None else np.concatenate([s["chr_id"] for s in samples])
        stream, _ = pgoracle.search(P, w["padded"], blob, off, np.concatenate([s["md"] for s in samples]), np.concatenate([s["pos"] for s in samples]),
                                    np.concatenate([s["isz"] for s in samples]), chr_id=chr_id, threads=os.cpu_count() or 1)
        exp = refhook.parse_stream(stream, seq.shape[0])
        bad = 0
        for e, g in zip(exp, got):
            if not (e["flip"] == g["flip"] and e["L"] == g["L"] and all(np.array_equal(e[key], g[key]) for key in ("close_raw", "close", "far"))):
                bad += 1
        log("parity (%s): %d reads from %d rank(s) checked against the oracle port, %d mismatches" % (spec["key"], seq.shape[0], world, bad))
        return {"checked": int(seq.shape[0]), "mismatches": int(bad), "ranks": world, "checker": "oracle/pg_oracle.c (pinned to the reference)"}
    except Exception as e:
        return {"checked": 0, "mismatches": None, "ranks": world, "error": repr(e)}


def main():
    a = parse_args()
    if a.cpu_worker:
        return cpu_worker(a.cpu_worker)
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

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the search path has no CPU fallback (use --impl reference for the CPU reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    want_cpu = rank == 0 and world == 1 and not a.no_cpu_baseline
    t0 = time.time()
    spec = spec_c3(a, world)
    w = make_workload(spec, rank, world, dev)
    log("headline workload generated in %.1fs" % (time.time() - t0))
    result = run_workload(a, spec, w, rank, local_rank, world, dev, a.steps, a.warmup, log, want_cpu)

    # ---- the other shapes of BASELINE.json, same measurement, fewer steps ----
    sel = [] if a.secondary == "none" else (["x5", "c2", "wg", "strong"] if a.secondary == "all" else a.secondary.split(","))
    sec_steps, sec_warm = max(2, min(a.steps, 5)), max(1, min(a.warmup, 3))
    secondary = []

    def run_secondary(sp, wl):
        try:
            r = run_workload(a, sp, wl, rank, local_rank, world, dev, sec_steps, sec_warm, log, want_cpu)
            if rank == 0:
                keep = {k: r[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "scaling", "e2e", "roofline", "parity", "gpu_launches", "clocks") if k in r}
                keep["config"] = r["config"]
                if "cpu_baseline" in r:
                    keep["cpu_baseline"] = r["cpu_baseline"]
                secondary.append(keep)
        except Exception as e:
            log("secondary %s failed: %r" % (sp["key"], e))
            if rank == 0:
                secondary.append({"config": {"workload": sp["workload"]}, "error": repr(e)})

    if "x5" in sel and a.max_range_index != 5:
        run_secondary(spec_c3(a, world, x=5), w)                 # the same reads, Pindel -x 5: windows up to 131 072 positions
    if "strong" in sel and world > 1:
        per = 50_000_000 // world                                 # BASELINE.json configs[3]: 50 M reads in total
        sp = spec_c3(a, world, reads=per, scaling="strong")
        sp["key"] = "c4_strong"
        run_secondary(sp, make_workload(sp, rank, world, dev) if per != a.reads else w)
    del w
    if "c2" in sel:
        sp = spec_c2(a)
        run_secondary(sp, make_workload(sp, rank, world, dev))
    if "wg" in sel:
        sp = spec_wg(a, world)
        t0 = time.time()
        try:
            wl = make_workload(sp, rank, world, dev)
            log("3 Gbp workload generated in %.1fs" % (time.time() - t0))
            run_secondary(sp, wl)
            del wl
        except Exception as e:
            log("3 Gbp workload failed: %r" % (e,))
            if rank == 0:
                secondary.append({"config": {"workload": sp["workload"]}, "error": repr(e)})
    if rank == 0:
        result["secondary"] = secondary
        emit(result)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def reference_arm(a, rank, world, log):
    """The reference's own CPU implementation of the path (GetCloseEnd + SearchFarEnds from oracle/_ref, built from the
    untouched sources), all host threads, same metric / config / generator output as the GPU arm's rank 0; each step is a
    bounded sample of that batch sized for a few seconds of OpenMP work (short steps under-read many-core hosts)."""
    if rank != 0:
        return
    try:
        import torch
        dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    except Exception:
        dev = None
    spec = spec_c3(a, max(1, world))
    if dev is not None and dev.type == "cuda":
        w = make_workload(spec, 0, max(1, world), dev)
        gen = "the GPU arm's generator (rank 0's batch)"
    else:
        # no CUDA device for the synthesis: the same read model from the numpy generator (same distribution, another stream)
        from pindel_b200 import synth
        ref_np = synth.make_reference(spec["ref_bases"], spec["ref_seed"])
        b = synth.make_reads(ref_np, min(a.reads, 2_000_000), a.read_len, seed=777, insert_size=INSERT)
        w = dict(padded=[synth.pad_reference(ref_np)], seq=b.seq, md=b.matched_d, pos=b.matched_rel_pos, isz=b.insert_size, chr_id=None)
        gen = "numpy generator (no CUDA device for the GPU arm's generator)"
    total = 150.0                                                 # seconds of CPU work for W + K steps
    per_step = max(2.0, min(6.0, total / max(1, a.steps + a.warmup)))
    ans = run_cpu_leg(w, spec, per_step, log, steps=a.steps, warmup=a.warmup)
    ms_step = ans["seconds"] * 1e3
    value = ans["reads"] / ans["seconds"]
    cfg = config_of(spec, max(1, world))
    sample = ("%d reads per step (bounded sample of the %d-read batch of %s), GetCloseEnd loop + SearchFarEnds wall time, %d OpenMP threads"
              % (ans["reads"], spec["reads_per_gpu_per_step"], gen, ans["cores"]))
    result = {"impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
              "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
              "config": cfg, "reads_per_step_sampled": ans["reads"], "host": "OpenMP, %d host threads (rank 0 only)" % ans["cores"],
              "cpu_baseline": {"value": value, "unit": "reads/s", "cores": ans["cores"], "kind": ans["kind"], "sample": sample},
              "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(result)


if __name__ == "__main__":
    main()
