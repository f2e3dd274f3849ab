"""Drop-in check at the file level: the UNMODIFIED reference host program (command line, Pindel-text / BAM-config input,
window loop, reporter) with its search seam functions and its seven SV drivers served by libpindel_pg.so
(oracle/ref_gpu_bridge.cpp, bound by the dynamic linker -- no reference source is patched) must write byte-identical
_D/_SI/_INV/_TD (and _LI/_BP/_CloseEndMapped) files to the ones the pure reference binary wrote for the same inputs
(tests/golden/colo_slice.*, sim_prefix.*).  Needs oracle/_ref/pindel_gpu_demo, which is built in the container where
/root/reference exists and travels to the GPU box with the snapshot."""
import os

import pytest

from tests import dropin_scenarios as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "oracle", "_ref", "pindel_gpu_demo")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(DEMO), reason="oracle/_ref/pindel_gpu_demo not built (needs /root/reference at build time)")]


def _env():
    # the freshly built product library, not a stale copy next to the demo
    return dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "pindel_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))


def test_reference_host_with_gpu_search_writes_identical_files(tmp_path):
    p = D.colo_slice(DEMO, tmp_path, _env())
    assert "(1 launches)" in p.stderr.splitlines()[-1]     # one window -> one classification launch for all seven drivers


def test_reference_host_bam_input_with_gpu_search(tmp_path):
    D.sim_prefix_bam(DEMO, tmp_path, _env())


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "pindel_ref")), reason="oracle/_ref/pindel_ref not built")
def test_interchromosomal_scenario_against_the_reference_binary(tmp_path):
    """-I true on generated inputs, the pure reference binary beside the drop-in: inter-chromosomal pair discovery and
    ModifyRP_InterChr / Summarize_InterChr on the device."""
    D.interchr_live(DEMO, tmp_path, _env())
