"""Randomised parity sweep on the GPU: random option sets (-x -a -m -H -u -e -E), mixed read lengths inside one batch
(so reads whose level count T is below the compiled TL take the cooperative path), several chromosomes, references with
N runs, random BreakDancer-style clusters -- each batch compared bit for bit with the oracle port (itself pinned to the
reference, tests/test_oracle_pinning.py)."""
import numpy as np
import pytest

from oracle import pgoracle, refhook
from pindel_b200 import api, runs as R, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def mixed_batch(refs, rng, n_per, lens):
    seqs, md, pos, isz, chr_id = [], [], [], [], []
    for c, ref in enumerate(refs):
        for L in lens:
            b = synth.make_reads(ref, n_per, int(L), seed=int(rng.integers(1, 1 << 30)), insert_size=int(rng.choice([200, 350, 500])),
                                 max_del=int(rng.choice([300, 3000])), n_rate=float(rng.choice([0.0, 0.002, 0.02])))
            for i in range(b.n):
                seqs.append(b.seq[i].tobytes())
            md.append(b.matched_d); pos.append(b.matched_rel_pos); isz.append(b.insert_size)
            chr_id.append(np.full(b.n, c, np.int32))
    order = rng.permutation(len(seqs))
    seqs = [seqs[i] for i in order]
    md, pos, isz, chr_id = (np.concatenate(x)[order] for x in (md, pos, isz, chr_id))
    off = np.zeros(len(seqs) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(s) for s in seqs])
    return np.frombuffer(b"".join(seqs), dtype=np.uint8), off, md, pos, isz, chr_id


def fuzz_case(seed, n_per=350):
    """-> (settings, padded chromosomes, blob, off, md, pos, isz, chr_id, regions (list of triples), region_off, lens)"""
    rng = np.random.default_rng(1000 + seed)
    settings = dict(max_range_index=int(rng.integers(0, 5)), additional_mismatch=int(rng.integers(1, 4)),
                    min_perfect_match=int(rng.integers(0, 8)), min_close=int(rng.integers(4, 14)),
                    max_mismatch_rate=float(rng.choice([0.0, 0.01, 0.02, 0.05, 0.1])),
                    seq_error_rate=float(rng.choice([0.005, 0.01, 0.02, 0.04])), sensitivity=float(rng.choice([0.9, 0.95, 0.99])))
    refs = [synth.make_reference(int(rng.integers(250_000, 400_000)), int(rng.integers(1, 1 << 30)),
                                 n_runs=[(int(rng.integers(20_000, 200_000)), int(rng.integers(1, 3000)))]) for _ in range(int(rng.integers(1, 4)))]
    padded = [synth.pad_reference(r) for r in refs]
    lens = rng.choice([40, 51, 64, 75, 100, 101, 150, 200, 251, 300], size=3, replace=False)
    blob, off, md, pos, isz, chr_id = mixed_batch(refs, rng, n_per, lens)
    n = len(off) - 1
    regions, roff = [], [0]
    for i in range(n):
        if rng.random() < 0.2:
            for _ in range(int(rng.integers(1, 4))):
                c = int(rng.integers(0, len(refs)))
                s = int(rng.integers(synth.SPACER, padded[c].size - synth.SPACER - 2000))
                regions.append((c, s, s + int(rng.integers(50, 1800))))
        roff.append(len(regions))
    return settings, padded, blob, off, md, pos, isz, chr_id, regions, np.array(roff, np.uint32), lens


def region_array(regions):
    rg = np.zeros(len(regions), dtype=R.REGION_DT)
    if regions:
        a = np.array(regions)
        rg["chr"], rg["start"], rg["end"] = a[:, 0], a[:, 1], a[:, 2]
    return rg


def oracle_answers(settings, padded, blob, off, md, pos, isz, chr_id, regions, roff):
    P = pgoracle.default_params(**settings)
    stream, _ = pgoracle.search(P, padded, blob, off, md, pos, isz, chr_id=chr_id,
                                regions=np.array(regions, dtype=np.uint32).reshape(-1, 3), region_off=roff, threads=8)
    return refhook.parse_stream(stream, len(off) - 1)


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_settings_and_mixed_lengths(seed):
    settings, padded, blob, off, md, pos, isz, chr_id, regions, roff, lens = fuzz_case(seed)
    pg = api.PatternGrowth(**settings)
    pg.set_reference(padded)
    out, runs = pg.search(seq_blob=blob, seq_off=off, matched_d=md, matched_rel_pos=pos, insert_size=isz, chr_id=chr_id,
                          regions=region_array(regions), region_off=roff)
    assert (out["flags"] & (R.F_TOOLONG | R.F_OVERFLOW)).sum() == 0
    got = R.per_read(out, runs, blob, off)
    H.compare(oracle_answers(settings, padded, blob, off, md, pos, isz, chr_id, regions, roff), got,
              "fuzz seed %d %s lens %s" % (seed, settings, lens.tolist()))
    pg.close()
