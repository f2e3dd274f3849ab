"""The CUDA warp program (pindel_b200/csrc/pg_warp.cuh) executed by the CPU lane emulator (tests/emul/), against the
reference's own answers (tests/golden/).  This is the host-side check of the device logic: same source, same
collectives, no GPU.  Every read of every committed dump of the reference goes through it (the emulator runs about 15 k
reads/s: fibers switched by a stack swap, one emulated warp per host thread)."""
import numpy as np
import pytest

from pindel_b200 import runs as R
from tests import fixtures as F, helpers as H
from tests.emul import emul


@pytest.fixture(scope="module")
def lib():
    return emul.build()


def run_emul(lib, case, lo=0, hi=None):
    prm = emul.make_params(lib, **{dict(seq_error_rate="seq_error_rate", sensitivity="sensitivity", max_mismatch_rate="max_mismatch_rate",
                                       additional_mismatch="add_mm", min_close="min_close", min_perfect_match="min_perfect",
                                       max_range_index="max_range_index")[k]: v for k, v in case.settings.items()})
    kw = case.kwargs(lo, hi)
    out, pool = emul.search(lib, prm, [case.padded], kw["seq_blob"], kw["seq_off"], kw["matched_d"], kw["matched_rel_pos"], kw["insert_size"])
    assert (out["flags"] & (R.F_TOOLONG | R.F_OVERFLOW)).sum() == 0
    return R.per_read(out, pool, kw["seq_blob"], kw["seq_off"])


@pytest.mark.parametrize("name,count", [("L100_default", 1200), ("L150_default", 1000), ("L75_x4_u05", 800),
                                        ("L100_a2_m5_H10_e03", 800), ("L250_default", 400)])
def test_warp_program_matches_reference_dump(lib, name, count):
    case = F.load_synth(name)
    assert case.n == count
    H.compare(case.expected, run_emul(lib, case, 0, count), "emulated warp program vs golden " + name)


def test_warp_program_edge_cases(lib):
    case = F.load_edge()
    H.compare(case.expected, run_emul(lib, case), "emulated warp program vs golden edge cases")


def test_warp_program_colo_slice(lib):
    case = F.load_colo_slice()
    assert case.n == 14077           # all of the slice, the reads nearest the centromere N run (last in the file) included
    H.compare(case.expected, run_emul(lib, case), "emulated warp program vs golden COLO slice")


def test_empty_batch(lib):
    case = F.load_edge()
    prm = emul.make_params(lib)
    out, pool = emul.search(lib, prm, [case.padded], np.zeros(0, np.uint8), np.zeros(1, np.uint32), np.zeros(0, np.uint8),
                            np.zeros(0, np.uint32), np.zeros(0, np.int16))
    assert out.size == 0 and pool.size == 0
