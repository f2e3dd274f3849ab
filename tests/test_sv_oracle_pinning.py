"""Pins oracle/sv_oracle.c (the CPU restatement of the seven SV drivers) to the reference:
against the committed dumps of the reference's own drivers (tests/golden/sv_*.npz) and, when oracle/_ref is built, against
the live reference on a fresh seeded batch."""
import os
import subprocess
import sys

import pytest

from oracle import refhook, svoracle
from tests import sv_helpers as S
from tests.golden.make_sv_golden import CASES


@pytest.mark.parametrize("name", list(CASES))
def test_sv_oracle_matches_reference_dumps(name):
    c = S.load_sv_case(name)
    svoracle.compare_dumps(c["golden"], S.oracle_dump(c), name)
    took = sum(len(v["vars"]) + len(v["future"]) for v in c["golden"].values())
    assert took > 100, "fixture exercises too few events"


LIVE = r"""
import sys, os, tempfile, numpy as np
sys.path.insert(0, %r)
from oracle import refhook, svoracle
from pindel_b200 import synth
from tests import sv_helpers as S
ref = synth.make_reference(500_000, 9201, n_runs=[(250_000, 1500)])
batch = synth.make_reads(ref, 2500, 110, seed=9202)
tmp = tempfile.mkdtemp(); fa = os.path.join(tmp, "r.fa")
refhook.write_fasta(fa, [("c", ref)])
h = refhook.RefHook(fa, ["-e", "0.02", "-v", "20"], workdir=tmp)
h.set_window(0, 0, ref.size)
blob, off = batch.blob()
s1, _ = h.search(blob, off, batch.matched_d, batch.matched_rel_pos, batch.insert_size, 0, True, 4)
per = refhook.parse_stream(s1, batch.n)
ci = S.classify_inputs(per, S.final_seqs(blob, off, per), batch.matched_d, batch.insert_size)
kw = {k: ci[k] for k in ("seq_blob", "seq_off", "matched_d", "insert_size", "rflags", "up_rows", "up_off", "n_close")}
for region, wend in (((0, ref.size), 350_000), ((100_000, 200_000), ref.size)):   # the region setting is sticky: widest first
    wp = S.window_params(h.chr_seq(0).size, wend)
    stream, _ = h.classify(**kw, window=(0, wend), num_boxes=wp["num_boxes"], box_size=wp["box_size"], region=region)
    P = svoracle.make_params(h.max_mismatch(), wend, wp["num_boxes"], wp["box_size"], region=((1, region[0]), (1, region[1])),
                             min_inv=h.setting_int("MIN_IndelSize_Inversion"), min_matched=h.setting_int("Min_Num_Matched_Bases"),
                             seq_error_rate=h.setting_double("Seq_Error_Rate"))
    svoracle.compare_dumps(svoracle.parse_dump(stream, len(ci["index"])), svoracle.classify(P, h.chr_seq(0).tobytes(), **kw), "live")
print("LIVE-OK")
"""


@pytest.mark.skipif(not refhook.available(), reason="oracle/_ref/libpindel_ref.so not built (needs /root/reference)")
def test_sv_oracle_matches_live_reference():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", LIVE % root], capture_output=True, text=True, timeout=600)
    assert "LIVE-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
