"""Discordant-pair discovery (first half): ReadInBamReads_RP_Discovery / build_record_RP_Discovery
(src/reader.cpp:418-480, 982-1097).  CPU part: the Python oracle and the device functions compiled for the host
(tests/emul/bam_emul.cpp) against the reference's own dumps (tests/golden/rpd_*.npz) and, when oracle/_ref is built, the
reference live.  The GPU part is tests/test_gpu_rp_discovery.py."""
import gzip
import os

import numpy as np
import pytest

from oracle import rpdoracle
from pindel_b200 import bam
from tests.emul import bam_emul
from tests.golden import make_rpd_golden as G

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def case_stream(case):
    if case == "sim_prefix":
        return gzip.decompress(open(os.path.join(GOLD, "sim_prefix.bam"), "rb").read())
    return bam.synthetic_stream(refs=G.SYNTH_REFS, **G.SYNTH)


def rows_of_product(stream, recs, rows, src, irows, isrc):
    """product output -> the 9-column layout of the reference dump"""
    out = []
    for lst, (rr, ss) in enumerate(((rows, src), (irows, isrc))):
        for r, i in zip(rr, ss):
            off = int(recs["off"][i]) + 4
            l_name = stream[off + 8]
            out.append((lst, int(r["pos_a"]), int(r["pos_b"]), int(r["da"]), int(r["db"]), int(r["read_length"]), int(r["insert_size"]),
                        int(recs["mapq"][i]), rpdoracle.fnv1a(stream[off + 32:off + 32 + l_name - 1])))
    return np.array(out, dtype=np.int64).reshape(-1, 9)


@pytest.fixture(scope="module")
def lib():
    return bam_emul.build()


@pytest.mark.parametrize("case", list(G.QUERIES))
def test_oracle_and_device_functions_match_reference_dumps(lib, case):
    z = np.load(os.path.join(GOLD, "rpd_%s.npz" % case))
    stream = case_stream(case)
    table = bam.serial_index(stream)
    recs, _ = bam_emul.index(lib, stream)
    nonempty = 0
    for k, (tid, beg, end, isz) in enumerate(G.QUERIES[case]):
        exp, scanned = z["rows_%d" % k], int(z["scanned_%d" % k])
        got, _, sc = rpdoracle.discovery(stream, table, tid, beg, end, isz)
        assert sc == scanned and np.array_equal(got, exp), ("oracle", case, k)
        rows, src, irows, isrc, sc2 = bam_emul.rp_discovery(lib, stream, recs, tid, beg, end, isz)
        assert sc2 == scanned and np.array_equal(rows_of_product(stream, recs, rows, src, irows, isrc), exp), ("device functions", case, k)
        nonempty += exp.shape[0] > 0
    assert nonempty >= 2


def test_oracle_matches_live_reference(tmp_path):
    from oracle import refhook
    if not refhook.available():
        pytest.skip("oracle/_ref not built")
    import subprocess
    import sys
    # the reference keeps process-global state: run it in a child, compare there
    code = """
import sys, numpy as np
sys.path.insert(0, %r)
from oracle import refhook, rpdoracle
from pindel_b200 import bam, bgzf
from tests.golden import make_rpd_golden as G
import tempfile
tmp = tempfile.mkdtemp()
refs = (("chrA", 300000), ("chrB", 200000))
stream = bam.synthetic_stream(5000, 77, refs=refs, discordant=0.4)
rng = np.random.default_rng(2)
fa = tmp + "/g.fa"
refhook.write_fasta(fa, [(n, rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), l)) for n, l in refs])
open(tmp + "/x.bam", "wb").write(bgzf.compress(stream)); open(tmp + "/x.bam.bai", "wb").write(b"BAI\\x01\\x00\\x00\\x00\\x00")
h = refhook.RefHook(fa, [], workdir=tmp)
table = bam.serial_index(stream)
for q in ((0, 0, 300000, 500), (0, 5000, 9000, 100), (0, 30000, 60000, 2000)):
    exp, scanned = h.rp_discovery(tmp + "/x.bam", *q)
    got, _, sc = rpdoracle.discovery(stream, table, *q)
    assert sc == scanned and np.array_equal(got, exp), q
print("LIVE-OK")
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert "LIVE-OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
