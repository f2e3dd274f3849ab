"""scripts/whole_program_sweep.py -- CPU-only, whole program: the PURE reference binary (oracle/_ref/pindel_ref, -T 1) beside the
reference host bound to this library's ABI over the device programs' host builds (oracle/_ref/pindel_emul_demo) under random
command lines -- search options (-x -a -m -e -E -u -d -H), reporter thresholds (-n -v -B -M -A), report switches (-r -t -l -k -s)
and the window size (-w, so that a chromosome is cut into several windows and reads are handed from window to window) -- on the
reference's own inputs: the simulated BAM (test/SmallTest/sim1chrVs2.bam, whole chromosome) and the COLO-829 text file (a
region of chr20).  Every output file must be byte-identical.  Needs /root/reference; not part of pytest.
    python scripts/whole_program_sweep.py [--rounds N] [--seed S] [--jobs J]"""
import argparse
import concurrent.futures as cf
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
SMALL = os.path.join(REF, "test", "SmallTest")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "pindel_ref")
DEMO_BIN = os.path.join(ROOT, "oracle", "_ref", "pindel_emul_demo")
ALL = ("D", "SI", "INV", "TD", "LI", "BP", "RP", "INT_final", "CloseEndMapped")


def draw(seed):
    rng = np.random.default_rng(seed)
    bam = bool(rng.integers(0, 2))
    a = ["-x", int(rng.integers(1, 5)), "-a", int(rng.integers(1, 3)), "-m", int(rng.integers(1, 6)), "-e", float(rng.choice([0.005, 0.01, 0.03])),
         "-E", float(rng.choice([0.9, 0.95, 0.99])), "-u", float(rng.choice([0.01, 0.02, 0.05, 0.1])), "-d", int(rng.choice([20, 30, 40])),
         "-H", int(rng.choice([4, 8, 12])), "-n", int(rng.choice([20, 50, 100])), "-v", int(rng.choice([20, 50, 200])), "-B", int(rng.choice([0, 100])),
         "-M", int(rng.choice([1, 2, 3, 5])), "-A", int(rng.choice([0, 10, 20, 30]))]
    for flag in ("-r", "-t", "-l", "-k", "-s"):
        a += [flag, "true" if rng.integers(0, 2) else "false"]
    if bam:
        a += ["-f", os.path.join(SMALL, "sim1chrVs2.fa"), "-i", os.path.join(SMALL, "sim1chrVs2.conf_for_test"), "-w", float(rng.choice([0.03, 0.07, 0.2, 5]))]
    else:
        lo = int(rng.integers(0, 23)) * 1_000_000            # (the -1 file holds the reads of 0-27 Mbp)
        a += ["-f", os.path.join(SMALL, "hs_ref_chr20.fa"), "-p", os.path.join(SMALL, "COLO-829_20-p_ok-1.txt"),
              "-c", "20:%d-%d" % (lo, lo + int(rng.choice([1, 2, 4])) * 1_000_000), "-w", float(rng.choice([0.5, 1, 5]))]
    return [str(x) for x in a]


DROPIN_THREADS = 1


def one(seed):
    args = draw(seed)
    tmp = tempfile.mkdtemp(prefix="pgsweep_")
    cwd = os.path.join(REF, "devtools")                     # the BAM configuration names its file relative to devtools/
    r = subprocess.run([REF_BIN] + args + ["-T", "1", "-o", os.path.join(tmp, "ref")], cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        return seed, args, "reference failed: " + (r.stdout + r.stderr)[-400:], 0
    p = subprocess.run([DEMO_BIN] + args + ["-T", str(DROPIN_THREADS), "-o", os.path.join(tmp, "got")], cwd=cwd, capture_output=True, text=True,
                       env=dict(os.environ, PG_EMUL_THREADS="3"))
    if p.returncode != 0:
        return seed, args, "drop-in failed: " + (p.stdout + p.stderr)[-400:], 0
    total = 0
    for suf in (ALL if DROPIN_THREADS == 1 else ("D", "SI", "INV", "TD", "LI", "BP")):     # (_RP: the host's own OpenMP loop orders its lines)
        a, b = (open(os.path.join(tmp, k + "_" + suf), "rb").read() if os.path.exists(os.path.join(tmp, k + "_" + suf)) else None for k in ("ref", "got"))
        if a != b:
            return seed, args, "_%s differs (%s vs %s bytes)" % (suf, a and len(a), b and len(b)), 0
        total += len(a or b"")
    return seed, args, None, total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=12)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--jobs", type=int, default=3)
    ap.add_argument("--dropin-threads", type=int, default=1, help="-T of the drop-in (the reference always runs with -T 1, the only "
                    "count at which its files are reproducible); above 1 the event files are compared, _RP is not")
    a = ap.parse_args()
    global DROPIN_THREADS
    DROPIN_THREADS = a.dropin_threads
    assert os.path.exists(REF_BIN) and os.path.exists(DEMO_BIN) and os.path.isdir(SMALL), "needs /root/reference and oracle/_ref"
    bad, nbytes, nonempty = [], 0, 0
    with cf.ThreadPoolExecutor(a.jobs) as ex:
        for seed, args, err, total in ex.map(one, range(a.seed, a.seed + a.rounds)):
            short = " ".join(x if not x.startswith("/") else os.path.basename(x) for x in args)
            if err:
                bad.append(seed)
                print("ROUND-FAILED seed %d: %s\n   %s" % (seed, err, short), flush=True)
            else:
                nbytes += total
                nonempty += total > 0
                print("round ok seed %d (%d bytes of output): %s" % (seed, total, short), flush=True)
    if bad:
        print("sweep FAILED: seeds %s" % bad)
        sys.exit(1)
    print("sweep clean: %d command lines (%d with events to report), %d bytes of output files identical to the pure reference binary's" % (a.rounds, nonempty, nbytes))


if __name__ == "__main__":
    main()
