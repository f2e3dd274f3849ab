"""A few rounds of scripts/search_hostile_campaign.py inside the suite: the search's warp program on the lane emulator with hostile
per-read fields (anchor positions anywhere in 32 bits, any insert size, arbitrary read bytes and lengths, clusters outside their
chromosome, arbitrary close ends) terminates, keeps its run offsets inside the pool and still answers the intact reads of the
batch.  The sanitizer run of the same campaign is scripts/sanitize_search_hostile.sh (profiles/r4_cpu_campaigns.txt)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))


def test_hostile_batches_terminate_and_stay_inside_their_buffers():
    import search_hostile_campaign as camp
    total, with_runs = camp.run(argparse.Namespace(rounds=12, reads=300, seed=11))
    assert total == 3600 and with_runs > 50


def test_damaged_bam_records_end_in_error_flags_not_in_a_crash():
    """scripts/bam_corruption_campaign.py in small: record index, pair discovery and split-read filter (host builds) on streams
    whose records are damaged behind an intact block_size chain; under the sanitizers: scripts/sanitize_bam_corruption.sh"""
    import bam_corruption_campaign as camp
    indexed, rows, cands, flagged = camp.run(argparse.Namespace(rounds=25, seed=21))
    assert indexed == 25 and rows > 1000 and cands > 1000 and flagged > 0
