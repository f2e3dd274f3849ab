"""The SV-classification device program (pindel_b200/csrc/pg_sv.cuh, body of pg_sv_kernel) compiled for the host: its
records, replayed by the host mirror (pindel_b200.sv.replay), must reproduce the reference's dumps bit for bit.  No GPU."""
import numpy as np
import pytest

from oracle import svoracle
from pindel_b200 import sv
from tests import sv_helpers as S
from tests.emul import sv_emul
from tests.golden.make_sv_golden import CASES


@pytest.fixture(scope="module")
def lib():
    return sv_emul.build()


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("raw_close", [False, True])
def test_classify_program_matches_reference_dumps(lib, name, raw_close):
    c = S.load_sv_case(name)
    b = S.product_batch(c, raw_close=raw_close)   # raw lists: the program applies CleanUniquePoints' filter itself
    rec = sv_emul.classify(lib, c["case"].padded, b, c["win"], seq_error_rate=c["seq_error_rate"], sensitivity=c["sensitivity"])
    got = S.dump_from_records(rec, c["seqs"], b["rflags"], driver_mask=c["win"]["driver_mask"])
    svoracle.compare_dumps(c["golden"], got, name)


def test_compress_ups_round_trip():
    from pindel_b200 import runs as R
    c = S.load_sv_case("synth_L100_default")
    for r in c["case"].expected[:300]:
        for key in ("close_raw", "close", "far"):
            rows = r[key]
            back = R.expand_runs(sv.compress_ups(rows))
            assert np.array_equal(back, rows), key
