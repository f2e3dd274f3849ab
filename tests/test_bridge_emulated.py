"""CPU test of the drop-in bridge's host logic (oracle/ref_gpu_bridge.cpp): the same bridge object, linked against the
CPU emulators of the device programs (tests/emul/abi_emul.cpp -> oracle/_ref/pindel_emul_demo) instead of
libpindel_pg.so, must reproduce the reference's output files.  What this covers without a GPU: marshalling of
SPLIT_READs into the ABI's SoA batches, the batched ReadBuffer::flush / ReadInRead bindings, replay of the classification
records per driver (Vars[box], FutureReads_SR, SPLIT_READ fields) and the hand-over to the reference's reporter.  The
GPU run of the same scenarios is tests/test_gpu_dropin.py."""
import os

import pytest

from tests import dropin_scenarios as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "oracle", "_ref", "pindel_emul_demo")

pytestmark = pytest.mark.skipif(not os.path.exists(DEMO), reason="oracle/_ref/pindel_emul_demo not built (needs /root/reference at build time)")


def test_text_input_scenario(tmp_path):
    p = D.colo_slice(DEMO, tmp_path, dict(os.environ), unbatched_too=False)
    assert "(1 launches)" in p.stderr.splitlines()[-1]


def test_bam_input_scenario(tmp_path):
    D.sim_prefix_bam(DEMO, tmp_path, dict(os.environ))


def test_bam_input_scenario_with_zlib_inflate(tmp_path):
    """the reader's own zlib path (PG_BRIDGE_ZLIB=1) instead of pg_bgzf_inflate: both must give the same files"""
    D.sim_prefix_bam(DEMO, tmp_path, dict(os.environ, PG_BRIDGE_ZLIB="1"))


def test_bam_input_scenario_with_the_split_read_filter_and_search_apart(tmp_path):
    """The default (test_bam_input_scenario) binds ReadInBamReads_SR to pg_bam_sr_close_end: filter and GetCloseEnd in one
    call.  PG_BRIDGE_SR_UNCHAINED=1 keeps them apart (pg_bam_sr_candidates, then the ReadBuffer's batched search) -- same files."""
    D.sim_prefix_bam(DEMO, tmp_path, dict(os.environ, PG_BRIDGE_SR_UNCHAINED="1"))


def test_bam_input_scenario_with_the_reference_reader(tmp_path):
    """PG_BRIDGE_HOST_SR=1: the reference's own ReadInBamReads_SR (per-record loop on the host) -- same files"""
    D.sim_prefix_bam(DEMO, tmp_path, dict(os.environ, PG_BRIDGE_HOST_SR="1"))


def test_bam_input_scenario_with_the_reference_mask(tmp_path):
    """The default binds BDData::loadRegion to pg_bd_load_region (clusters looked up inside pg_far_end).  PG_BRIDGE_HOST_BD=1:
    the reference's own position-by-position mask and getCorrespondingSearchWindowCluster per read -- same files"""
    D.sim_prefix_bam(DEMO, tmp_path, dict(os.environ, PG_BRIDGE_HOST_BD="1"))


def test_bam_input_scenario_over_a_device_group(tmp_path):
    """PG_BRIDGE_DEVICES=0,0: the BreakDancer region is loaded into every context of the group and each looks its own reads up"""
    D.sim_prefix_bam(DEMO, tmp_path, dict(os.environ, PG_BRIDGE_DEVICES="0,0", PG_BRIDGE_NO_RESIDENT="1"))


def test_text_input_scenario_over_a_device_group(tmp_path):
    """PG_BRIDGE_DEVICES=3: the two batched seams go through pg_group_close_end / pg_group_far_end (reads dealt to three
    contexts, records merged with global run offsets), the classification takes the host arrays -- same files"""
    D.colo_slice(DEMO, tmp_path, dict(os.environ, PG_BRIDGE_DEVICES="3", PG_BRIDGE_NO_RESIDENT="1"), unbatched_too=False)


def test_text_input_scenario_with_the_host_array_classification(tmp_path):
    """pg_classify (everything uploaded again) instead of pg_classify_resident: the path a host takes when the far-end
    batch is no longer on the device; and the reference's own bubblesortReads / markDuplicates instead of pg_box_sort."""
    D.colo_slice(DEMO, tmp_path, dict(os.environ, PG_BRIDGE_NO_RESIDENT="1", PG_BRIDGE_HOST_SORT="1"), unbatched_too=False)


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "pindel_ref")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/pindel_ref not built")
def test_dense_di_and_inversion_events_against_the_reference_binary(tmp_path):
    """row N4, the rest: SortOutputDI and OutputSorter::DoSortAndOutputInversions supplied by the bridge around pg_box_sort_kind"""
    D.dense_events_live(DEMO, tmp_path, dict(os.environ))


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/pindel_ref not built")
def test_dense_di_and_inversion_events_with_the_reference_reporters(tmp_path):
    D.dense_events_live(DEMO, tmp_path, dict(os.environ, PG_BRIDGE_HOST_SORT="1", PG_BRIDGE_NO_RESIDENT="1"))


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/pindel_ref not built")
def test_interchromosomal_scenario_against_the_reference_binary(tmp_path):
    D.interchr_live(DEMO, tmp_path, dict(os.environ))
