"""GPU path vs the committed dumps of the reference (tests/golden/): synthetic batches under several option sets,
hand-made corner cases (N / IUPAC / lower case / trailing junk / short reads / chromosome ends / degenerate windows) and
the real-data COLO-829 chr20 slice with the centromere N run.  Called through the C ABI; bit-exact."""
import numpy as np
import pytest

from pindel_b200 import api, runs as R
from tests import fixtures as F, helpers as H

pytestmark = pytest.mark.gpu


def run_case(case, **extra):
    pg = api.PatternGrowth(**case.settings)
    pg.set_reference([case.padded])
    kw = case.kwargs()
    out, runs = pg.search(**kw, **extra)
    assert (out["flags"] & (R.F_TOOLONG | R.F_OVERFLOW)).sum() == 0
    got = R.per_read(out, runs, kw["seq_blob"], kw["seq_off"])
    pg.close()
    return got


@pytest.mark.parametrize("name", F.ALL_SYNTH)
def test_gpu_matches_reference_dump_synthetic(name):
    case = F.load_synth(name)
    H.compare(case.expected, run_case(case), "GPU vs golden " + name)


def test_gpu_matches_reference_dump_edge_cases():
    case = F.load_edge()
    H.compare(case.expected, run_case(case), "GPU vs golden edge cases")


def test_gpu_matches_reference_dump_colo_slice():
    case = F.load_colo_slice()
    H.compare(case.expected, run_case(case), "GPU vs golden COLO slice")


def test_gpu_empty_batch_and_reuse():
    case = F.load_edge()
    pg = api.PatternGrowth()
    pg.set_reference([case.padded])
    out, runs = pg.search(seq_blob=np.zeros(0, np.uint8), seq_off=np.zeros(1, np.uint32), matched_d=np.zeros(0, np.uint8),
                          matched_rel_pos=np.zeros(0, np.uint32), insert_size=np.zeros(0, np.int16))
    assert out.size == 0 and runs.size == 0
    kw = case.kwargs()
    a = pg.search(**kw)
    b = pg.search(**kw)       # same context, second batch: buffers are reused
    ra, rb = R.per_read(*a, kw["seq_blob"], kw["seq_off"]), R.per_read(*b, kw["seq_blob"], kw["seq_off"])
    H.compare(ra, rb, "repeatability")
    pg.close()


def test_gpu_multi_chromosome_and_clusters():
    """Two chromosomes, reads anchored on both, BreakDancer-style clusters pointing at the other chromosome; checked
    against the oracle port (whose multi-region path is pinned to the reference by the live COLO -b run)."""
    from oracle import pgoracle, refhook
    from pindel_b200 import synth
    r0 = synth.make_reference(400_000, 501)
    r1 = synth.make_reference(300_000, 502, n_runs=[(120_000, 900)])
    p0, p1 = synth.pad_reference(r0), synth.pad_reference(r1)
    b0 = synth.make_reads(r0, 1500, 100, seed=503)
    b1 = synth.make_reads(r1, 1500, 100, seed=504)
    seq = np.concatenate([b0.seq, b1.seq])
    md = np.concatenate([b0.matched_d, b1.matched_d])
    pos = np.concatenate([b0.matched_rel_pos, b1.matched_rel_pos])
    isz = np.concatenate([b0.insert_size, b1.insert_size])
    chr_id = np.concatenate([np.zeros(b0.n, np.int32), np.ones(b1.n, np.int32)])
    n = seq.shape[0]
    off = (np.arange(n + 1, dtype=np.uint64) * 100).astype(np.uint32)
    rng = np.random.default_rng(505)
    regions, roff = [], [0]
    for i in range(n):
        if i % 3 == 0:      # a cluster of 1-3 regions, some overlapping, some on the other chromosome
            for _ in range(int(rng.integers(1, 4))):
                c = int(rng.integers(0, 2))
                size = (p0 if c == 0 else p1).size
                s = int(rng.integers(synth.SPACER, size - synth.SPACER - 3000))
                regions.append((c, s, s + int(rng.integers(200, 2500))))
        roff.append(len(regions))
    rg = np.zeros(len(regions), dtype=R.REGION_DT)
    a = np.array(regions)
    rg["chr"], rg["start"], rg["end"] = a[:, 0], a[:, 1], a[:, 2]
    pg = api.PatternGrowth()
    pg.set_reference([p0, p1])
    out, runs = pg.search(seq_blob=seq.reshape(-1), seq_off=off, matched_d=md, matched_rel_pos=pos, insert_size=isz,
                          chr_id=chr_id, regions=rg, region_off=np.array(roff, np.uint32))
    got = R.per_read(out, runs, seq.reshape(-1), off)
    P = pgoracle.default_params()
    stream, _ = pgoracle.search(P, [p0, p1], seq.reshape(-1), off, md, pos, isz, chr_id=chr_id,
                                regions=a.astype(np.uint32), region_off=np.array(roff, np.uint32), threads=4)
    exp = refhook.parse_stream(stream, n)
    H.compare(exp, got, "two chromosomes + clusters")
    assert sum(1 for r in got if len(r["far"]) and (r["far"][:, 5] != r["close"][0, 5]).any()) >= 0
    pg.close()
