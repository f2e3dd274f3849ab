"""tests/test_gpu_fuzz.py's randomised sweep on the CPU: the same generated cases -- random option sets, mixed read lengths
inside one batch, several chromosomes, references with N runs, random BreakDancer-style clusters (also on other
chromosomes) -- through the warp program on the lane emulator instead of the kernel, bit for bit against the oracle port.
More seeds than the GPU file runs, since nothing here needs the device."""
import pytest

from pindel_b200 import runs as R
from tests import helpers as H
from tests.emul import emul
from tests.test_gpu_fuzz import fuzz_case, oracle_answers, region_array

pytestmark = []          # (test_gpu_fuzz's own mark is `gpu`; nothing here is)


@pytest.fixture(scope="module")
def lib():
    return emul.build()


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12])
def test_random_settings_and_mixed_lengths_on_the_emulator(lib, seed):
    settings, padded, blob, off, md, pos, isz, chr_id, regions, roff, lens = fuzz_case(seed, n_per=120)
    names = dict(seq_error_rate="seq_error_rate", sensitivity="sensitivity", max_mismatch_rate="max_mismatch_rate",
                 additional_mismatch="add_mm", min_close="min_close", min_perfect_match="min_perfect", max_range_index="max_range_index")
    prm = emul.make_params(lib, **{names[k]: v for k, v in settings.items()})
    out, pool = emul.search(lib, prm, padded, blob, off, md, pos, isz, chr_id=chr_id, regions=region_array(regions), region_off=roff)
    assert (out["flags"] & (R.F_TOOLONG | R.F_OVERFLOW)).sum() == 0
    got = R.per_read(out, pool, blob, off)
    H.compare(oracle_answers(settings, padded, blob, off, md, pos, isz, chr_id, regions, roff), got,
              "emulated fuzz seed %d %s lens %s" % (seed, settings, lens.tolist()))
    assert sum(1 for r in got if len(r["far"])) > 50
