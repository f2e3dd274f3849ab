"""scripts/cpu_fuzz_campaign_r2.py -- longer randomized differential runs for the stages added in round 2, CPU only:

  * BreakDancer mask: the closed form of pg_bd.cuh (host build) against the REFERENCE's own BDData::loadRegion +
    getCorrespondingSearchWindowCluster (oracle/_ref/libpindel_ref.so, pgref_bd_clusters) on fresh event lists;
  * box ordering of every reporter: the kernel's formulation (host build, lanes written out) against oracle/boxsort_oracle.c;
  * the DI / inversion reporters end to end: the PURE REFERENCE BINARY (oracle/_ref/pindel_ref) beside the reference host on the
    emulated ABI (oracle/_ref/pindel_emul_demo, reporters supplied around pg_box_sort_kind) on inputs dense in those events.

Not part of the pytest suite (minutes of CPU).   python scripts/cpu_fuzz_campaign_r2.py [--rounds N] [--seed S]"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=40)
    ap.add_argument("--seed", type=int, default=2026)
    a = ap.parse_args()
    from oracle import bdoracle, boxsortoracle as B, refhook
    from tests import bd_helpers as BH, boxsort_helpers as H, dense_events
    from tests.emul import bd_emul, boxsort_emul
    assert refhook.available(), "needs oracle/_ref (build in the container that has /root/reference)"
    rng = np.random.default_rng(a.seed)
    tmp = tempfile.mkdtemp()
    fa = os.path.join(tmp, "g.fa")
    refhook.write_fasta(fa, [(n, rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), l)) for n, l in zip(BH.NAMES, BH.LENS)])
    hook = refhook.RefHook(fa, [], workdir=tmp)
    bd_lib, box_lib = bd_emul.build(), boxsort_emul.build()
    n_q = n_clusters = n_boxes = 0
    for r in range(a.rounds):
        # ---- BreakDancer mask against the live reference ----
        chr_id, window, ev, q = BH.case(100_000 + 1000 * (a.seed - 2026) + r, dense=bool(r % 3 == 0), n_queries=800)
        exp = bdoracle.reference_clusters(hook, ev, chr_id, window[0], window[1], q)
        ev_sorted = np.array(bdoracle.sort_on_first(bdoracle.normalise(ev), BH.NAMES), dtype=np.int64).reshape(-1, 6)
        got = bd_emul.clusters(bd_lib, ev_sorted, BH.NAMES, chr_id, window[0], window[1], q)
        we = window[1] + 3000
        for k, (e, g) in enumerate(zip(exp, got)):
            if q[k] >= we:
                continue                      # the reference reads mask entries it never wrote there
            assert e == g, ("bd", r, k, int(q[k]), e, g)
            n_clusters += bool(e)
        n_q += q.size
        # ---- every reporter's loops: kernel formulation against the restatement ----
        for kind in range(5):
            s = 200_000 + 10 * r + kind + 1000 * (a.seed - 2026)
            if kind <= 1:
                reads, box_off, idx = B.make_boxes(int(rng.integers(1, 150)), s, max_reads=int(rng.integers(3, 130)), heavy=int(rng.integers(0, 2)))
            else:
                reads, box_off, idx = B.make_boxes_kind(kind, int(rng.integers(1, 150)), s, max_reads=int(rng.integers(3, 130)), heavy=int(rng.integers(0, 2)))
            o_idx, o_unique, o_event = B.sort_boxes_kind(kind, reads, box_off, idx)
            order, non, ev_flags = boxsort_emul.sort_boxes_kind(box_lib, kind, reads[idx], box_off, True)
            p_idx, p_unique = H.apply(reads, box_off, idx, order, non)
            assert np.array_equal(p_idx, o_idx) and np.array_equal(p_unique, o_unique) and np.array_equal(ev_flags, o_event), ("box", r, kind)
            n_boxes += len(box_off) - 1
        print("round %d ok" % r, flush=True)
    # ---- the DI / inversion reporters at the file level, several seeds ----
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "pindel_ref")
    demo = os.path.join(ROOT, "oracle", "_ref", "pindel_emul_demo")
    n_events = 0
    for k in range(max(3, a.rounds // 5)):
        d = tempfile.mkdtemp()
        ref, text, n = dense_events.make(seed=300 + 97 * k + (a.seed - 2026))
        refhook.write_fasta(os.path.join(d, "g.fa"), [("dense", ref)])
        open(os.path.join(d, "g.fa.fai"), "w").write("dense\t%d\t7\t60\t61\n" % ref.size)
        open(os.path.join(d, "reads.txt"), "w").write(text)
        args = ["-f", os.path.join(d, "g.fa"), "-p", os.path.join(d, "reads.txt"), "-k", "-s", "-l", "-T", "1"]
        for exe, pre in ((ref_bin, "ref"), (demo, "dropin")):
            p = subprocess.run([exe] + args + ["-o", os.path.join(d, pre)], capture_output=True, text=True, timeout=1800)
            assert p.returncode == 0, p.stderr[-500:]
        for suf in ("D", "SI", "INV", "TD", "LI", "BP", "CloseEndMapped"):
            x, y = open(os.path.join(d, "ref_" + suf), "rb").read(), open(os.path.join(d, "dropin_" + suf), "rb").read()
            assert x == y, ("dense events", k, suf)
            if suf in ("D", "INV"):
                n_events += sum(1 for line in x.split(b"\n") if line.startswith(b"####"))
        print("dense-event input %d ok (%d reads)" % (k, n), flush=True)
    print("campaign clean: %d rounds; %d BreakDancer lookups (%d non-empty clusters) equal to the reference; %d boxes over the 5 reporters equal to the "
          "restatement; %d D / INV events in files identical to the reference binary's" % (a.rounds, n_q, n_clusters, n_boxes, n_events))


if __name__ == "__main__":
    main()
