"""Host-side parameter derivation of the library (pg_params.h via the emulator build, which compiles the very same
header): the mismatch-budget table must equal the reference's createProbTable bit for bit, the CheckMismatches
threshold must reproduce the float comparison of src/searcher.cpp:366,383."""
import numpy as np
import pytest

from oracle import pgoracle
from tests.emul import emul


@pytest.fixture(scope="module")
def lib():
    return emul.build()


def tables(lib, **kw):
    blob = emul.make_params(lib, **kw)
    og, ot = lib.pgemu_sizes(5), lib.pgemu_sizes(6)          # offsetof(PgParams, g / thr)
    g = np.frombuffer(blob.tobytes()[og:og + 512], dtype=np.uint8)
    thr = np.frombuffer(blob.tobytes()[ot:ot + 1024], dtype=np.uint16)
    return g, thr


@pytest.mark.parametrize("e,E", [(0.01, 0.95), (0.03, 0.95), (0.005, 0.99), (0.05, 0.9), (0.02, 0.999)])
def test_mismatch_table_equals_oracle(lib, e, E):
    g, _ = tables(lib, seq_error_rate=e, sensitivity=E)
    assert np.array_equal(g[:500].astype(np.uint32), pgoracle.prob_table(e, E))


@pytest.mark.parametrize("u", [0.02, 0.05, 0.0, 0.1, 0.013])
def test_threshold_is_the_float_comparison(lib, u):
    _, thr = tables(lib, max_mismatch_rate=u)
    for L in (0, 1, 36, 75, 100, 149, 150, 151, 250, 499):
        max_allowed = np.float32(np.float64(L) * np.float64(u))
        n = int(thr[L])
        assert np.float32(n) >= max_allowed
        assert n == 0 or np.float32(n - 1) < max_allowed
