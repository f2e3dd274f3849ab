"""Golden vectors of the discordant-pair discovery (row N2, first half) made by the REFERENCE ITSELF: its own
ReadInBamReads_RP_Discovery (src/reader.cpp:418-480, 982-1097) through oracle/ref_hook.cpp, on the committed BAM fixture and
on a seeded synthetic two-chromosome BAM.  Run in the container where /root/reference exists (one process per fasta):
    python -m tests.golden.make_rpd_golden
The .npz holds, per query, the rows the reference pushed and the number of records its iterator delivered."""
from __future__ import annotations

import gzip
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SYNTH = dict(n_records=6000, seed=21, discordant=0.25)
SYNTH_REFS = (("chr1", 400_000), ("chr2", 300_000))
# case -> list of (chr, start, end, insert_size)
QUERIES = {
    "sim_prefix": [(0, 0, 10_000_000, 500), (0, 150_000, 250_000, 500), (0, 100_000, 100_050, 200), (0, 390_000, 400_000, 500)],
    "synthetic": [(0, 0, 400_000, 500), (1, 0, 300_000, 250), (0, 20_000, 30_000, 250), (0, 1000, 2000, 500), (0, 46_000, 48_000, 500)],
}


def synthetic_bam() -> bytes:
    from pindel_b200 import bam, bgzf
    return bgzf.compress(bam.synthetic_stream(refs=SYNTH_REFS, **SYNTH))


def stage(case: str, tmp: str):
    """-> (fasta path, bam path) of a case, written under tmp"""
    from oracle import refhook
    if case == "sim_prefix":
        fa = os.path.join(tmp, "sim_prefix.fa")
        open(fa, "wb").write(gzip.open(os.path.join(HERE, "sim_prefix.fa.gz")).read())
        shutil.copy(os.path.join(HERE, "sim_prefix.fa.fai"), fa + ".fai")
        bam_path = os.path.join(tmp, "sim_prefix.bam")
        shutil.copy(os.path.join(HERE, "sim_prefix.bam"), bam_path)
    else:
        rng = np.random.default_rng(1)
        fa = os.path.join(tmp, "synthetic.fa")
        refhook.write_fasta(fa, [(n, rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), l)) for n, l in SYNTH_REFS])
        bam_path = os.path.join(tmp, "synthetic.bam")
        open(bam_path, "wb").write(synthetic_bam())
    open(bam_path + ".bai", "wb").write(b"BAI\x01\x00\x00\x00\x00")
    return fa, bam_path


def run_case(case: str):
    from oracle import refhook
    tmp = tempfile.mkdtemp()
    fa, bam_path = stage(case, tmp)
    h = refhook.RefHook(fa, [], workdir=tmp)
    out = {}
    for k, (c, a, b, isz) in enumerate(QUERIES[case]):
        rows, scanned = h.rp_discovery(bam_path, c, a, b, isz)
        out["rows_%d" % k] = rows
        out["scanned_%d" % k] = np.int64(scanned)
        print(case, (c, a, b, isz), rows.shape[0], "rows,", int((rows[:, 0] == 1).sum()) if rows.size else 0, "inter-chromosomal,", scanned, "scanned")
    np.savez_compressed(os.path.join(HERE, "rpd_%s.npz" % case), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for case in QUERIES:       # the reference keeps process-global state: one process per genome
            subprocess.check_call([sys.executable, "-m", "tests.golden.make_rpd_golden", case], cwd=ROOT)
