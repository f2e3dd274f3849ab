"""The far rings' threshold filter (pg_filter_tile in pindel_b200/csrc/pg_warp.cuh) on the CPU lane emulator: the warp
program against the oracle port at larger -x, where the rings are long enough for the filter, on references built to
stress it -- near-identical repeats and low-complexity stretches inside the far windows (large thresholds, ties, many
well-matching seeds), N runs inside the rings, reads that start with N or carry N / IUPAC codes in their first 32 bases,
chromosomes so short that every window is clipped.  Bit-exact or it fails."""
import numpy as np
import pytest

from oracle import pgoracle, refhook
from pindel_b200 import runs as R, synth
from tests import helpers as H
from tests.emul import emul


@pytest.fixture(scope="module")
def lib():
    return emul.build()


def repetitive_reference(n, seed):
    rng = np.random.default_rng(seed)
    ref = synth.make_reference(n, seed)
    # near-identical copies of a few blocks, a couple of kb apart (inside the far windows of the reads planted there)
    for _ in range(n // 20_000):
        ln = int(rng.integers(200, 3000))
        a = int(rng.integers(5000, n - 20_000))
        b = a + int(rng.integers(ln, 9000))
        blk = ref[a:a + ln].copy()
        mut = rng.random(ln) < rng.choice([0.0, 0.01, 0.05])
        blk[mut] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(mut.sum()))]
        ref[b:b + ln] = blk
    # low-complexity stretches and short N runs
    for _ in range(n // 40_000):
        a = int(rng.integers(5000, n - 5000))
        unit = np.frombuffer(rng.choice([b"A", b"AC", b"AGT", b"TTTA"]), dtype=np.uint8)
        ln = int(rng.integers(40, 400))
        ref[a:a + ln] = np.resize(unit, ln)
    for _ in range(n // 60_000):
        a = int(rng.integers(5000, n - 5000))
        ref[a:a + int(rng.integers(1, 300))] = ord("N")
    return ref


def run_both(lib, padded, batch, filter_min_pos=None, **settings):
    names = dict(seq_error_rate="seq_error_rate", sensitivity="sensitivity", max_mismatch_rate="max_mismatch_rate",
                 additional_mismatch="add_mm", min_close="min_close", min_perfect_match="min_perfect", max_range_index="max_range_index")
    prm = emul.make_params(lib, **{names[k]: v for k, v in settings.items()})
    if filter_min_pos is not None:
        prm.view(np.uint32)[4] = filter_min_pos      # PgParams::filter_min_pos: rings from this length on are filtered
    blob, off = batch.blob()
    emul.filter_stats(lib)
    out, pool = emul.search(lib, prm, [padded], blob, off, batch.matched_d, batch.matched_rel_pos, batch.insert_size)
    assert (out["flags"] & (R.F_TOOLONG | R.F_OVERFLOW)).sum() == 0
    seeds, survivors, tiles = emul.filter_stats(lib)
    # the filter ran, and it did filter: most seeds of the rings are dropped before the per-seed evaluation
    assert tiles > 0 and seeds > 0 and survivors < seeds
    print("filter: %d tiles, %d seeds in, %d survivors (%.2f %%)" % (tiles, seeds, survivors, 100.0 * survivors / seeds))
    got = R.per_read(out, pool, blob, off)
    exp = H.oracle_results([padded], batch, threads=8, **settings)
    return exp, got


@pytest.mark.parametrize("x,L,n,seed", [(3, 100, 600, 11), (4, 150, 400, 12), (5, 75, 240, 13), (6, 100, 120, 14)])
def test_filter_rings_on_repetitive_reference(lib, x, L, n, seed):
    ref = repetitive_reference(400_000, seed)
    padded = synth.pad_reference(ref)
    batch = synth.make_reads(ref, n, L, seed=seed + 100, max_del=6000, n_rate=0.004)
    exp, got = run_both(lib, padded, batch, max_range_index=x)
    H.compare(exp, got, "filter rings, -x %d, L %d" % (x, L))
    assert sum(1 for r in got if len(r["far"])) > n // 4


def test_filter_with_special_characters_in_the_first_word(lib):
    """N and IUPAC codes among the read's first 32 characters are left out of the filter's step table (they can only
    lower a mismatch count); the survivors' exact evaluation applies Matches()' rules."""
    ref = repetitive_reference(300_000, 21)
    padded = synth.pad_reference(ref)
    batch = synth.make_reads(ref, 40, 100, seed=22, max_del=3000)
    rng = np.random.default_rng(23)
    for i in range(batch.n):
        for _ in range(int(rng.integers(1, 4))):
            batch.seq[i, int(rng.integers(1, 30))] = rng.choice(np.frombuffer(b"NNNRYn", dtype=np.uint8))
        if i % 7 == 0:
            batch.seq[i, 0] = ord("N")
    exp, got = run_both(lib, padded, batch, max_range_index=4)
    H.compare(exp, got, "filter rings, special characters")


def test_filter_on_a_short_chromosome(lib):
    """Every far window is clipped by the chromosome ends (src/pindel.cpp:1033-1042), so rings are lopsided or empty."""
    ref = repetitive_reference(30_000, 31)
    padded = synth.pad_reference(ref)
    batch = synth.make_reads(ref, 40, 100, seed=32, max_del=2000, insert_size=300)
    exp, got = run_both(lib, padded, batch, max_range_index=5, additional_mismatch=2)
    H.compare(exp, got, "filter rings, short chromosome")


def test_filter_on_short_rings(lib):
    """The library filters rings of 2048 positions and more (the fourth window on); with the knob turned down the second and
    third windows of Pindel's default -x 2 go through it as well -- thresholds from 128 positions, most seeds survive."""
    ref = repetitive_reference(300_000, 41)
    padded = synth.pad_reference(ref)
    batch = synth.make_reads(ref, 80, 100, seed=42, max_del=900, n_rate=0.003)
    exp, got = run_both(lib, padded, batch, filter_min_pos=64, max_range_index=2)
    H.compare(exp, got, "filter on the short rings of -x 2")
