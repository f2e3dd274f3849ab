"""Randomised three-way check of the split-read filter: the reference's ReadInBamReads_SR live through the hook (when oracle/_ref
is built), the Python oracle, and the device functions on the host (tests/emul/sr_emul.cpp), on streams that visit the corners of
the BAM format (tests/sr_fuzz_helpers.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import sroracle
from pindel_b200 import bam
from tests import sr_fuzz_helpers as F
from tests.emul import bam_emul, sr_emul
from tests.test_sr_emulator import rows_of_product

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_device_functions_match_oracle(seed):
    lib, bl = sr_emul.build(), bam_emul.build()
    stream = F.fuzz_stream(seed)
    table = bam.serial_index(stream)
    recs, _ = bam_emul.index(bl, stream)
    total = 0
    for tid, beg, end, isz, min_q in F.QUERIES:
        exp_c, exp_r = sroracle.candidates(stream, table, tid, beg, end, isz, F.REFS[tid][1], min_anchor_quality=min_q)
        cands, seqs, refs, _, err = sr_emul.candidates(lib, stream, recs, sr_emul.query(tid, beg, end, isz, F.REFS[tid][1], min_q), seed, 0,
                                                       0 if seed % 2 else 0x3FF << 20)
        assert err == 0
        c, r = rows_of_product(stream, recs, cands, seqs, refs, isz)
        assert np.array_equal(c, exp_c) and np.array_equal(r, exp_r), (seed, tid, beg, end, isz, min_q)
        total += exp_c.shape[0] + exp_r.shape[0]
    assert total > 1000


def test_oracle_matches_live_reference():
    from oracle import refhook
    if not refhook.available():
        pytest.skip("oracle/_ref not built")
    # the reference keeps process-global state: run it in a child, compare there
    code = """
import sys, tempfile, numpy as np
sys.path.insert(0, %r)
from oracle import refhook, sroracle
from pindel_b200 import bam, bgzf
from tests import sr_fuzz_helpers as F
tmp = tempfile.mkdtemp()
rng = np.random.default_rng(2)
fa = tmp + "/g.fa"
refhook.write_fasta(fa, [(n, rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), l)) for n, l in F.REFS])
h = refhook.RefHook(fa, [], workdir=tmp)
for seed in (11, 12, 13):
    stream = F.fuzz_stream(seed)
    path = tmp + "/x%%d.bam" %% seed
    open(path, "wb").write(bgzf.compress(stream)); open(path + ".bai", "wb").write(b"BAI\\x01\\x00\\x00\\x00\\x00")
    table = bam.serial_index(stream)
    for tid, beg, end, isz, min_q in F.QUERIES:
        cand, refs = h.sr_candidates(path, tid, beg, end, isz, min_q)
        oc, orf = sroracle.candidates(stream, table, tid, beg, end, isz, F.REFS[tid][1], min_anchor_quality=min_q)
        assert np.array_equal(cand, oc) and np.array_equal(refs, orf), (seed, tid, beg, end, isz, min_q, cand.shape, oc.shape)
print("LIVE-OK")
""" % ROOT
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert "LIVE-OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
