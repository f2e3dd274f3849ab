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


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "pindel_ref")), reason="oracle/_ref/pindel_ref not built")
def test_dense_di_and_inversion_events_against_the_reference_binary(tmp_path):
    """Row N4, the rest: the DI and inversion reporters (inlined ordering / duplicate scan / event cut) supplied around
    pg_box_sort_kind, on an input dense in those events, the pure reference binary beside the drop-in."""
    D.dense_events_live(DEMO, tmp_path, _env())


def test_reference_host_over_a_device_group(tmp_path):
    """PG_BRIDGE_DEVICES: the two batched seams fan out inside the library (pg_group_close_end / pg_group_far_end).  Three
    contexts on device 0 on a one-GPU box, the real devices when there are more -- same files."""
    import torch
    nd = torch.cuda.device_count()
    D.colo_slice(DEMO, tmp_path, dict(_env(), PG_BRIDGE_DEVICES=str(nd) if nd > 1 else "0,0,0", PG_BRIDGE_NO_RESIDENT="1"), unbatched_too=False)


def test_reference_host_bam_input_with_the_reference_reader(tmp_path):
    """PG_BRIDGE_HOST_SR=1: the reference's own ReadInBamReads_SR instead of the device filter + search -- same files"""
    D.sim_prefix_bam(DEMO, tmp_path, dict(_env(), PG_BRIDGE_HOST_SR="1"))
