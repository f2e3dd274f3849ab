"""scripts/sv_bench.py -- measurement of the SV-classification stage on one GPU, with the
reference's own CPU drivers timed beside it.

    python scripts/sv_bench.py [--reads N] [--len L] [--cpu-sample K]

Workload: the bench.py read model (synthetic reference, D/SI/INV/TD mix) searched on the GPU (pg_search); the reads with a
close end, their sequences after the flips and the search kernel's own run lists go to pg_classify (host arrays in, sorted
records out).  Reported: end-to-end reads/s of the pg_classify call (H2D + kernel + D2H + record sort), the kernel's
CUDA-event time, records per read, and -- when oracle/_ref is present -- the wall time of the reference's seven drivers
(SearchVariant::Search x2, searchIndels, searchTandemDuplications(NT), searchInversions(NT); reporters stubbed out) on a
sample of the same reads (one thread: the reference runs them serially, src/pindel.cpp:1140-1164).
Prints one JSON line."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--len", type=int, default=150)
    ap.add_argument("--ref", type=int, default=62_435_964)
    ap.add_argument("--cpu-sample", type=int, default=200_000)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch
    from pindel_b200 import api, runs as R, sv, synth
    dev = torch.device("cuda", 0)
    ref = synth.make_reference(a.ref, 20260926)
    padded = synth.pad_reference(ref)
    seq_t, md_t, pos_t, isz_t, _ = synth.make_reads_torch(torch.from_numpy(ref).to(dev), a.reads, a.len, 4242)
    seq, md, pos, isz = seq_t.cpu().numpy(), md_t.cpu().numpy(), pos_t.cpu().numpy().astype(np.uint32), isz_t.cpu().numpy()
    L, n = a.len, a.reads
    off = (np.arange(n + 1, dtype=np.uint64) * L).astype(np.uint32)
    pg = api.PatternGrowth()
    pg.set_reference([padded])
    out, runs = pg.search(seq_blob=seq.reshape(-1), seq_off=off, matched_d=md, matched_rel_pos=pos, insert_size=isz)
    keep = np.nonzero(out["n_close_runs"] > 0)[0]          # the readers drop reads without a close end (src/reader.cpp:262-288)
    m = keep.size
    flips = ((out["flags"][keep] >> 8) & 3) & 1
    fin = seq[keep].copy()
    comp = np.zeros(256, np.uint8)
    for x, y in zip(b"ACGTN", b"TGCAN"):
        comp[x] = y
    fl = np.nonzero(flips)[0]
    fin[fl] = comp[fin[fl][:, ::-1]]
    cnt = (out["n_close_runs"][keep].astype(np.int64) + out["n_far_runs"][keep])
    run_off = np.zeros(m + 1, dtype=np.uint32)
    run_off[1:] = np.cumsum(cnt)
    src = np.repeat(out["run_off"][keep].astype(np.int64) - run_off[:-1], cnt) + np.arange(int(run_off[-1]))
    batch = dict(seq_blob=fin.reshape(-1), seq_off=(np.arange(m + 1, dtype=np.uint64) * L).astype(np.uint32), matched_d=md[keep],
                 insert_size=isz[keep], rflags=np.zeros(m, np.uint8), run_off=run_off,
                 n_close_runs=out["n_close_runs"][keep].astype(np.uint16), runs=runs[src])
    win = sv.window(int(padded.size), int(ref.size), region=(0, int(ref.size)))
    rec_buf = np.empty(2 * m + 1024, dtype=sv.RECORD_DT)
    rec = pg.classify(batch, win, records=rec_buf)           # warm-up (allocations)
    wall, kern = [], []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        rec = pg.classify(batch, win, records=rec_buf)
        wall.append(time.perf_counter() - t0)
        kern.append(pg.classify_last_kernel_ms)
    h2d = sum(int(np.asarray(v).nbytes) for v in batch.values())
    res = {"stage": "SV classification (7 drivers, one launch)", "reads": int(m), "read_len": L, "runs_per_read": float(run_off[-1]) / m,
           "records": int(rec.size), "records_by_outcome": np.bincount(rec["outcome"], minlength=4).tolist(),
           "e2e_reads_per_s": m / float(np.median(wall)), "e2e_ms": 1e3 * float(np.median(wall)), "kernel_ms": float(np.median(kern)),
           "kernel_ms_note": "CUDA events around the whole device stage: memsets + pg_sv_kernel + scan + order",
           "kernel_reads_per_s": m / (float(np.median(kern)) * 1e-3), "h2d_bytes": h2d, "d2h_bytes": int(rec.nbytes)}
    # the same stage on the batch the search left on the device (pg_classify_resident): no second upload; all n reads,
    # the ones without a close end fall out on the device
    pg.search(seq_blob=seq.reshape(-1), seq_off=off, matched_d=md, matched_rel_pos=pos, insert_size=isz)
    rec_r = pg.classify_resident(win, records=rec_buf)
    assert rec_r.size == rec.size, (rec_r.size, rec.size)
    wall_r, kern_r = [], []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        rec_r = pg.classify_resident(win, records=rec_buf)
        wall_r.append(time.perf_counter() - t0)
        kern_r.append(pg.classify_last_kernel_ms)
    res["resident"] = {"reads_in_batch": int(n), "e2e_ms": 1e3 * float(np.median(wall_r)), "device_ms": float(np.median(kern_r)),
                       "e2e_reads_per_s": n / float(np.median(wall_r)), "h2d_bytes": 4168, "d2h_bytes": int(rec_r.nbytes)}
    from oracle import refhook
    if refhook.available() and a.cpu_sample > 0:
        import tempfile
        k = min(a.cpu_sample, m)
        rows = R.expand_runs(batch["runs"][:int(run_off[k])])
        per_run = batch["runs"]["count"][:int(run_off[k])].astype(np.int64)
        ups_per_run_cum = np.concatenate(([0], np.cumsum(per_run)))
        up_off = ups_per_run_cum[run_off[:k + 1].astype(np.int64)].astype(np.uint32)
        close_end = ups_per_run_cum[(run_off[:k].astype(np.int64) + batch["n_close_runs"][:k])]
        # CleanUniquePoints on the close lists, vectorised: keep rows whose (dir, strand, terminal) equal the list's last row
        term = np.where(rows[:, 2] == 0, rows[:, 1] - rows[:, 0], rows[:, 1] + rows[:, 0]) & 0xFFFFFFFF
        rid = np.repeat(np.arange(k), np.diff(up_off.astype(np.int64)))
        is_close = np.arange(rows.shape[0]) < close_end[rid]
        last = close_end - 1
        keep_row = ~is_close | ((term == term[last][rid]) & (rows[:, 2] == rows[last, 2][rid]) & (rows[:, 3] == rows[last, 3][rid]))
        rows2 = rows[keep_row]
        up_off2 = np.concatenate(([0], np.cumsum(np.bincount(rid[keep_row], minlength=k)))).astype(np.uint32)
        n_close2 = np.bincount(rid[keep_row & is_close], minlength=k).astype(np.uint32)
        tmp = tempfile.mkdtemp()
        fa = os.path.join(tmp, "ref.fa")
        refhook.write_fasta(fa, [("c", ref)])
        h = refhook.RefHook(fa, [], workdir=tmp)
        csz = h.chr_seq(0).size
        box = csz // 30000 or 1
        stream, secs = h.classify(batch["seq_blob"][:k * L], batch["seq_off"][:k + 1], batch["matched_d"][:k], batch["insert_size"][:k],
                                  batch["rflags"][:k], rows2[:, :5].astype(np.int64).astype(np.uint32).view(np.int32) if False else
                                  (rows2[:, :5] & 0xFFFFFFFF).astype(np.uint32).view(np.int32), up_off2, n_close2,
                                  window=(0, int(ref.size)), num_boxes=csz * 2 // box + 1, box_size=box, region=(0, int(ref.size)))
        res["cpu_reference"] = {"reads": int(k), "seconds_by_driver": [round(x, 4) for x in secs], "seconds": float(sum(secs)),
                                "reads_per_s": k / float(sum(secs)), "threads": 1,
                                "note": "the reference's own driver loops (reporters stubbed), oracle/_ref/libpindel_ref.so"}
        res["speedup_e2e_vs_cpu"] = res["e2e_reads_per_s"] / res["cpu_reference"]["reads_per_s"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
