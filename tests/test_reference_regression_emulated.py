"""The reference's OWN regression suite and demo, whole program, on the CPU: the unmodified reference host bound to this
library's ABI (oracle/ref_gpu_bridge.cpp) with the device programs compiled for the host (tests/emul/abi_emul.cpp ->
oracle/_ref/pindel_emul_demo), at the inputs' full size:

  * devtools/run_regressiontests.sh:5   pindel -f hs_ref_chr20.fa -p COLO-829_20-p_ok-1.txt -k -s -l -T 1    (62 Mbp, 189 580 reads)
  * devtools/run_regressiontests.sh:7   pindel -f sim1chrVs2.fa -i sim1chrVs2.conf_for_test -k -s -l -T 1    (BAM input)
    both against the reference's committed files, devtools/gold_standard/ -- every one of them, as its `diff -q` does;
  * demo/RUNME (BASELINE.json configs[0])   pindel -i simulated_config.txt -f simulated_reference.fa -c ALL  (three BAMs)
  * src/update_target_files.sh:10-15        the whole COLO-829 file, -c 20:0-1,500,000 -x 5 -B 100 -u 0.05 -b <BreakDancer file>
    both against the SHA-256 of the files the pure reference binary writes (tests/golden/reference_whole_program_sha256.json,
    made by tests/golden/make_whole_program_golden.py; the committed test/SmallTest/TargetOutput/ is from an older Pindel);
  * the simulated BAM input with an external BreakDancer file (-b), the pure reference binary beside the drop-in.

Every stage of SURVEY.md section 8 runs in these: close and far ends (the warp program on the lane emulator), the SV drivers,
pair discovery and clustering, the BreakDancer mask, inflate + record index + split-read filter, the box ordering of all five
reporters.  The inputs live in /root/reference, which exists only in the build container: skipped elsewhere (the GPU box runs
the committed slices of the same scenarios, tests/test_gpu_dropin.py; the full-size GPU runs are profiles/r1_colo_gold_dropin.txt
and profiles/r1y_bam_dropin.txt)."""
import hashlib
import json
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO_BIN = os.path.join(ROOT, "oracle", "_ref", "pindel_emul_demo")
REF = "/root/reference"
GOLD = os.path.join(REF, "devtools", "gold_standard")
ALL = ("D", "SI", "INV", "TD", "LI", "BP", "RP", "INT_final", "CloseEndMapped")

pytestmark = pytest.mark.skipif(not (os.path.exists(DEMO_BIN) and os.path.isdir(GOLD)),
                                reason="needs /root/reference (gold standard + inputs) and oracle/_ref/pindel_emul_demo")


def _run(args, cwd, prefix):
    p = subprocess.run([DEMO_BIN] + args + ["-o", prefix, "-T", "1"], cwd=cwd, capture_output=True, text=True, timeout=1800)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    m = re.search(r"(\d+) GetCloseEnd calls and (\d+) far-end reads served by libpindel_pg", p.stderr)
    assert m and int(m.group(1)) > 0 and int(m.group(2)) > 0, p.stderr[-600:]
    m = re.search(r"(\d+) reads classified by the SV drivers on the device", p.stderr)
    assert m and int(m.group(1)) > 0, p.stderr[-600:]
    m = re.search(r"(\d+) boxes ordered on the device \((\d+) launches\), (\d+) by the reference", p.stderr)
    assert m and int(m.group(1)) > 0 and int(m.group(3)) == 0, p.stderr[-600:]
    return p


def _against_gold(tmp_path, name):
    n = 0
    for suf in ALL:
        exp = open(os.path.join(GOLD, "%s_%s" % (name, suf)), "rb").read()
        got = (tmp_path / ("%s_%s" % (name, suf))).read_bytes()
        assert got == exp, "%s_%s differs from devtools/gold_standard (%d vs %d bytes)" % (name, suf, len(got), len(exp))
        n += len(exp)
    return n


def test_colo_regression_against_the_gold_standard(tmp_path):
    small = os.path.join(REF, "test", "SmallTest")
    p = _run(["-f", os.path.join(small, "hs_ref_chr20.fa"), "-p", os.path.join(small, "COLO-829_20-p_ok-1.txt"), "-k", "-s", "-l"],
             os.path.join(REF, "devtools"), str(tmp_path / "colo_test.out"))
    assert _against_gold(tmp_path, "colo_test.out") > 1_000_000
    m = re.search(r"(\d+) GetCloseEnd calls", p.stderr)
    assert int(m.group(1)) == 189580                       # every read of the input went through the batched seam
    m = re.search(r"(\d+) boxes of the DI / inversion reporters ordered and cut into events there", p.stderr)
    assert m and int(m.group(1)) > 1000


def test_simulated_bam_regression_against_the_gold_standard(tmp_path):
    small = os.path.join(REF, "test", "SmallTest")
    # the configuration file names the BAM relative to devtools/ (the regression script's working directory)
    p = _run(["-f", os.path.join(small, "sim1chrVs2.fa"), "-i", os.path.join(small, "sim1chrVs2.conf_for_test"), "-k", "-s", "-l"],
             os.path.join(REF, "devtools"), str(tmp_path / "simulated_test.out"))
    assert _against_gold(tmp_path, "simulated_test.out") > 100_000
    m = re.search(r"(\d+) read pairs clustered on the device", p.stderr)
    assert m and int(m.group(1)) > 10_000
    m = re.search(r"(\d+) BAM bytes inflated and (\d+) records indexed on the device, (\d+) discordant pairs found there", p.stderr)
    assert m and int(m.group(1)) > 0 and int(m.group(2)) > 0 and int(m.group(3)) > 0, p.stderr[-600:]
    m = re.search(r"(\d+) split-read candidates from the device filter, (\d+) of them searched where the filter left them", p.stderr)
    assert m and int(m.group(1)) > 0 and m.group(1) == m.group(2), p.stderr[-600:]


def _against_hashes(tmp_path, name):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_whole_program_sha256.json")))[name]
    for suf, g in gold["files"].items():
        data = (tmp_path / ("%s_%s" % (name, suf))).read_bytes()
        assert len(data) == g["bytes"] and hashlib.sha256(data).hexdigest() == g["sha256"], "%s_%s differs from the reference binary's file" % (name, suf)
    return sum(g["bytes"] for g in gold["files"].values())


def test_demo_runme_against_the_pure_reference_binary(tmp_path):
    _run(["-i", "simulated_config.txt", "-f", "simulated_reference.fa", "-c", "ALL"], os.path.join(REF, "demo"), str(tmp_path / "demo_runme"))
    assert _against_hashes(tmp_path, "demo_runme") > 3_000_000


def test_target_file_command_with_wide_far_rings(tmp_path):
    """src/update_target_files.sh:10-15 (`colousingbd15`): the whole COLO-829 text file, region 20:0-1,500,000, -x 5 (far rings
    up to 131 k positions: the bit-sliced filter), -u 0.05, -B 100, and a BreakDancer file.  The reference copies a -b file's
    events into its working list only inside UpdateBD, which runs for BAM input alone (src/pindel.cpp:1838-1846,
    src/bddata.cpp:649) -- with text input the file is read and then ignored, so `colowobd15` (the same command without -b) has
    the same golden; the drop-in follows: its region holds no events."""
    small = os.path.join(REF, "test", "SmallTest")
    p = _run(["-p", os.path.join(small, "COLO-829_20-p_ok.txt"), "-f", os.path.join(small, "hs_ref_chr20.fa"), "-x", "5", "-B", "100", "-l", "-k",
              "-u", "0.05", "-c", "20:0-1,500,000", "-b", os.path.join(small, "COLO-829.20.BreakDancer.sv")], small, str(tmp_path / "colousingbd15"))
    assert _against_hashes(tmp_path, "colousingbd15") > 3_000_000
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_whole_program_sha256.json")))
    assert gold["colousingbd15"]["files"] == gold["colowobd15"]["files"]
    assert re.search(r"\(0 events\) loaded into the library", p.stderr), p.stderr[-600:]


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "pindel_ref")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/pindel_ref not built")
def test_bam_input_with_an_external_breakdancer_file(tmp_path):
    """-b with BAM input, where the reference does use it: 1500 seeded events over the simulated chromosome join the ones the
    pair discovery adds, BDData::loadRegion (bound to pg_bd_load_region) masks them and every read's cluster is looked up
    inside pg_far_end.  No committed golden: the pure reference binary runs beside the drop-in; and the events do change the
    calls -- _D, _TD and _LI differ from the regression run without -b."""
    import numpy as np
    small = os.path.join(REF, "test", "SmallTest")
    name, length = open(os.path.join(small, "sim1chrVs2.fa.fai")).read().split()[:2]
    rng = np.random.default_rng(5)
    with open(tmp_path / "ext.sv", "w") as f:
        f.write("#Chr1\tPos1\tOrientation1\tChr2\tPos2\tOrientation2\tType\tSize\tScore\tnum_Reads\n")
        for _ in range(1500):
            a, b = (int(x) for x in rng.integers(1000, int(length) - 1000, 2))
            f.write("%s\t%d\t2+2-\t%s\t%d\t2+2-\tDEL\t%d\t50\t4\tx\t0.1\n" % (name, a, name, b, abs(a - b)))
    args = ["-f", os.path.join(small, "sim1chrVs2.fa"), "-i", os.path.join(small, "sim1chrVs2.conf_for_test"), "-k", "-s", "-l", "-b", str(tmp_path / "ext.sv")]
    cwd = os.path.join(REF, "devtools")
    r = subprocess.run([REF_BIN] + args + ["-o", str(tmp_path / "ref"), "-T", "1"], cwd=cwd, capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    p = _run(args, cwd, str(tmp_path / "got"))
    for suf in ALL:
        assert (tmp_path / ("got_" + suf)).read_bytes() == (tmp_path / ("ref_" + suf)).read_bytes(), "_%s differs from the reference binary's file" % suf
    m = re.search(r"(\d+) BreakDancer regions \((\d+) events\) loaded into the library, (\d+) cluster lookups", p.stderr)
    assert m and int(m.group(2)) > 2500 and int(m.group(3)) > 10_000, p.stderr[-600:]
    changed = [suf for suf in ALL if (tmp_path / ("ref_" + suf)).read_bytes() != open(os.path.join(GOLD, "simulated_test.out_" + suf), "rb").read()]
    assert {"D", "TD", "LI"} <= set(changed), changed
