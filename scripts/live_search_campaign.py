"""scripts/live_search_campaign.py -- CPU-only, three-way: the search kernel's warp program (pg_warp.cuh on the lane emulator of
tests/emul/), the oracle port (oracle/pg_oracle.c) and the REFERENCE ITSELF (oracle/_ref/libpindel_ref.so: its own GetCloseEnd
loop + SearchFarEnds, initialised with the round's command-line options) on the same seeded reads.  Every round draws an option
set (-x -a -m -H -u -e -E as Pindel's command line takes them), a read length and a reference with repeats; the reference keeps
process-global settings, so each round is its own process.  Needs oracle/_ref (the container that has /root/reference); not part
of pytest.
    python scripts/live_search_campaign.py [--rounds N] [--reads M] [--seed S] [--jobs J]"""
import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one_round(seed, n_reads):
    from oracle import pgoracle, refhook
    from pindel_b200 import runs as R, synth
    from tests import helpers as H
    from tests.emul import emul
    rng = np.random.default_rng(seed)
    st = dict(max_range_index=int(rng.integers(0, 5)), add_mm=int(rng.integers(1, 4)), min_perfect=int(rng.integers(0, 8)),
              min_close=int(rng.integers(4, 14)), max_mismatch_rate=float(rng.choice([0.0, 0.01, 0.02, 0.05, 0.1])),
              seq_error_rate=float(rng.choice([0.005, 0.01, 0.02, 0.04])), sensitivity=float(rng.choice([0.9, 0.95, 0.99])))
    L = int(rng.choice([40, 64, 75, 100, 101, 150, 200, 250]))
    ref = synth.make_reference(int(rng.integers(250_000, 400_000)), int(rng.integers(1, 1 << 30)),
                               n_runs=[(int(rng.integers(20_000, 200_000)), int(rng.integers(1, 3000)))])
    batch = synth.make_reads(ref, n_reads, L, seed=int(rng.integers(1, 1 << 30)), max_del=int(rng.choice([200, 1000, 20000])))
    blob, off = batch.blob()
    tmp = tempfile.mkdtemp(prefix="pglive_")
    fa = os.path.join(tmp, "r.fa")
    refhook.write_fasta(fa, [("c", ref)])
    args = ["-x", st["max_range_index"], "-a", st["add_mm"], "-m", st["min_perfect"], "-H", st["min_close"], "-u", st["max_mismatch_rate"],
            "-e", st["seq_error_rate"], "-E", st["sensitivity"]]
    h = refhook.RefHook(fa, args, workdir=tmp)
    h.set_window(0, 0, ref.size)
    live, _ = h.search(blob, off, batch.matched_d, batch.matched_rel_pos, batch.insert_size, 0, True, 2)
    padded = h.chr_seq(0)
    P = pgoracle.default_params(max_range_index=st["max_range_index"], additional_mismatch=st["add_mm"], min_perfect_match=st["min_perfect"],
                                min_close=st["min_close"], max_mismatch_rate=st["max_mismatch_rate"], seq_error_rate=st["seq_error_rate"],
                                sensitivity=st["sensitivity"])
    assert np.array_equal(np.array(P.max_mismatch[:]), h.max_mismatch()), "mismatch table differs from the reference's g_maxMismatch"
    port, _ = pgoracle.search(P, [padded], blob, off, batch.matched_d, batch.matched_rel_pos, batch.insert_size, threads=2)
    assert np.array_equal(live, port), "port and live reference disagree: %s L=%d" % (st, L)
    lib = emul.build()
    prm = emul.make_params(lib, **st)
    out, pool = emul.search(lib, prm, [padded], blob, off, batch.matched_d, batch.matched_rel_pos, batch.insert_size)
    assert (out["flags"] & (R.F_TOOLONG | R.F_OVERFLOW)).sum() == 0
    H.compare(refhook.parse_stream(live, batch.n), R.per_read(out, pool, blob, off), "seed %d L=%d %s" % (seed, L, st))
    print("ROUND-OK seed=%d L=%d reads=%d %s" % (seed, L, batch.n, st), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=20)
    ap.add_argument("--reads", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=3001)
    ap.add_argument("--jobs", type=int, default=4)
    ap.add_argument("--worker", type=int, default=-1)
    a = ap.parse_args()
    if a.worker >= 0:
        one_round(a.worker, a.reads)
        return
    from oracle import refhook
    assert refhook.available(), "needs oracle/_ref (build in the container that has /root/reference)"
    env = dict(os.environ, PG_EMUL_THREADS="2", OMP_NUM_THREADS="2")
    seeds = [a.seed + k for k in range(a.rounds)]
    running, done, failed = [], 0, []
    while seeds or running:
        while seeds and len(running) < a.jobs:
            s = seeds.pop(0)
            running.append((s, subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", str(s), "--reads", str(a.reads)],
                                                stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)))
        s, p = running.pop(0)
        out = p.communicate()[0]
        if p.returncode == 0 and "ROUND-OK" in out:
            done += 1
            print([ln for ln in out.splitlines() if ln.startswith("ROUND-OK")][-1], flush=True)
        else:
            failed.append(s)
            print("ROUND-FAILED seed=%d\n%s" % (s, out[-3000:]), flush=True)
    if failed:
        print("campaign FAILED: seeds %s" % failed)
        sys.exit(1)
    print("campaign clean: %d rounds, %d reads: warp program (emulated) == oracle port == live reference" % (done, done * a.reads))


if __name__ == "__main__":
    main()
