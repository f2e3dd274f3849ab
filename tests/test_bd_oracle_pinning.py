"""oracle/bdoracle.py (restatement of BDData::loadRegion / createRegionCluster / getCorrespondingSearchWindowCluster) against the
reference's own functions, driven through the hook library on event lists that reach the sweep's corners."""
import os

import numpy as np
import pytest

from oracle import bdoracle, refhook
from tests import bd_helpers as B

pytestmark = pytest.mark.skipif(not refhook.available(), reason="oracle/_ref/libpindel_ref.so not built (needs /root/reference at build time)")


@pytest.fixture(scope="module")
def hook(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("bd")
    rng = np.random.default_rng(5)
    fa = os.path.join(tmp, "g.fa")
    refhook.write_fasta(fa, [(n, rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), l)) for n, l in zip(B.NAMES, B.LENS)])
    return refhook.RefHook(fa, [], workdir=str(tmp))


@pytest.mark.parametrize("seed,dense", [(1, False), (2, False), (3, True), (4, False), (5, True), (6, False)])
def test_oracle_equals_reference(hook, seed, dense):
    rng = np.random.default_rng(100 + seed)
    chr_id = int(rng.integers(0, 3))
    ws = int(rng.choice([B.SPACER, B.SPACER + 500, B.SPACER + 40_000, 1000]))
    window = (ws, ws + int(rng.choice([20_000, 60_000])))
    ev = B.random_events(rng, int(rng.choice([0, 3, 40, 200])) if not dense else 150, chr_id, window, dense)
    # events anchored on other chromosomes sort before / after: they must not leak into the selection
    other = B.random_events(rng, 20, (chr_id + 1) % 3, window)
    ev = np.concatenate([ev, other])
    q = B.queries(rng, window, 400)
    exp = bdoracle.reference_clusters(hook, ev, chr_id, window[0], window[1], q)
    got = bdoracle.clusters(ev, B.NAMES, chr_id, window[0], window[1], q)
    n_nonempty = 0
    for k, (e, g) in enumerate(zip(exp, got)):
        if g is None:
            continue                          # the reference reads one element past its mask there
        assert e == g, (seed, k, int(q[k]), e, g)
        n_nonempty += bool(e)
    if ev.shape[0] > 60:
        assert n_nonempty > 10
