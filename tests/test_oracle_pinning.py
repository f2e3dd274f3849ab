"""Pins the CPU restatement (oracle/pg_oracle.c) to the reference: against the committed dumps of the reference
(tests/golden/) everywhere, and against the live reference library (oracle/_ref) where it was built."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pgoracle, refhook
from tests import fixtures as F, helpers as H


def run_port(case, threads=4):
    P = pgoracle.default_params(**case.settings)
    stream, _ = pgoracle.search(P, [case.padded], case.blob, case.off, case.md, case.pos, case.isz, threads=threads)
    return refhook.parse_stream(stream, case.n)


@pytest.mark.parametrize("name", F.ALL_SYNTH)
def test_port_matches_reference_dump_synthetic(name):
    case = F.load_synth(name)
    H.compare(case.expected, run_port(case), "oracle port vs golden " + name)


def test_port_matches_reference_dump_edge_cases():
    case = F.load_edge()
    H.compare(case.expected, run_port(case), "oracle port vs golden edge cases")


def test_port_matches_reference_dump_colo_slice():
    case = F.load_colo_slice()
    got = run_port(case)
    H.compare(case.expected, got, "oracle port vs golden COLO slice")
    assert sum(1 for r in got if len(r["far"])) > 5000     # the fixture is not degenerate


def test_prob_table_known_values():
    # default table (e=0.01+0.001, E=0.95)
    t = pgoracle.prob_table()
    assert list(t[:4]) == [0, 0, 0, 0] and t[4] == 1
    assert set(t[5:33]) == {2} and set(t[33:75]) == {3} and set(t[75:126]) == {4}
    assert set(t[126:181]) == {5} and set(t[181:239]) == {6}


LIVE = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from oracle import refhook, pgoracle
from pindel_b200 import synth
import tempfile, os
ref = synth.make_reference(500_000, 9001, n_runs=[(250_000, 1500)])
batch = synth.make_reads(ref, 1500, 120, seed=9002)
tmp = tempfile.mkdtemp(); fa = os.path.join(tmp, "r.fa")
refhook.write_fasta(fa, [("c", ref)])
h = refhook.RefHook(fa, ["-x", "3"], workdir=tmp)
h.set_window(0, 0, ref.size)
blob, off = batch.blob()
s1, _ = h.search(blob, off, batch.matched_d, batch.matched_rel_pos, batch.insert_size, 0, True, 4)
P = pgoracle.default_params(max_range_index=3)
s2, _ = pgoracle.search(P, [h.chr_seq(0)], blob, off, batch.matched_d, batch.matched_rel_pos, batch.insert_size, threads=4)
assert np.array_equal(np.array(P.max_mismatch[:]), h.max_mismatch())
assert np.array_equal(s1, s2), "port and live reference disagree"
print("LIVE-OK")
"""


@pytest.mark.skipif(not refhook.available(), reason="oracle/_ref/libpindel_ref.so not built (needs /root/reference)")
def test_port_matches_live_reference():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", LIVE % root], capture_output=True, text=True, timeout=600)
    assert "LIVE-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
