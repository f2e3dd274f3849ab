"""GPU parity of the reader's split-read filter (rest of row N3): pg_bgzf_inflate + pg_bam_index +
pg_bam_sr_candidates through the C ABI against the reference's own dumps (tests/golden/sr_*.npz) and the oracle, and the
whole-program drop-in with ReadInBamReads_SR bound to it.

The file sorts last on purpose: this stage was finished after the round's GPU minutes were spent, so it is CPU-verified only
(tests/test_sr_emulator.py runs the same __host__ __device__ functions, tests/test_bridge_emulated.py the same bridge code) and its
first run on the hardware is the round-end run of this file.  A failure here concerns this stage alone."""
import os

import numpy as np
import pytest

from oracle import sroracle
from pindel_b200 import api, bam, bgzf
from tests import dropin_scenarios as D
from tests.golden import make_sr_golden as G
from tests.test_sr_emulator import case_stream, named_stream, rows_of_product

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
DEMO = os.path.join(ROOT, "oracle", "_ref", "pindel_gpu_demo")


def _diff(got, exp):
    """a message that says where two row tables part (this file's first hardware run is unattended)"""
    if got.shape != exp.shape:
        k = min(got.shape[0], exp.shape[0])
        bad = np.nonzero((got[:k] != exp[:k]).any(axis=1))[0]
        return "shape %s vs %s; first differing row %s" % (got.shape, exp.shape, (int(bad[0]), got[bad[0]].tolist(), exp[bad[0]].tolist()) if bad.size else "none in the common prefix")
    bad = np.nonzero((got != exp).any(axis=1))[0]
    return "%d of %d rows differ; first: row %d %s vs %s" % (bad.size, got.shape[0], int(bad[0]), got[bad[0]].tolist(), exp[bad[0]].tolist()) if bad.size else "equal"


@pytest.fixture(scope="module")
def pg():
    p = api.PatternGrowth()
    yield p
    p.close()


def _run_setup(pg):
    from tests.test_bam_index_emulator import run_expectation
    stream = G.paired_stream(6000, 19, refs=G.SYNTH_REFS[:1])
    table = bam.serial_index(stream)
    block = 0x8000
    n_blocks = -(-len(stream) // block)
    lo, hi = 5 * block, (n_blocks - 4) * block
    exp, first, tail = run_expectation(table, len(stream), lo, hi)
    run = stream[lo:hi]
    pg.bgzf_inflate(bgzf.compress(run, block=block), to_host=False)
    recs, _ = pg.bam_index_run(first, 1)
    inside = table[(table["off"] >= lo + first) & (table["off"] < lo + tail)]
    return stream, table, run, recs, int(inside["pos"][50]), int(inside["pos"][-50])


def test_index_of_a_run_of_blocks(pg):
    """pg_bam_index_run (CPU-verified only so far, like the filter below): a window's chunk of BGZF blocks instead of the whole
    file -- no header in front, the last record cut off -- gives the serial walk's records of that stretch, and the discordant-pair
    discovery (kernels that have run on the hardware) on it gives what it gives on the whole file"""
    from tests.test_bam_index_emulator import run_expectation, same
    from oracle import rpdoracle
    from tests.test_rp_discovery import rows_of_product as rpd_rows
    stream = G.paired_stream(6000, 19, refs=G.SYNTH_REFS[:1])
    table = bam.serial_index(stream)
    block = 0x8000
    n_blocks = -(-len(stream) // block)
    rng = np.random.default_rng(3)
    for _ in range(6):
        b0 = int(rng.integers(1, n_blocks - 2))
        b1 = int(rng.integers(b0 + 1, n_blocks + 1))
        lo, hi = b0 * block, min(b1 * block, len(stream))
        exp, first, tail = run_expectation(table, len(stream), lo, hi)
        pg.bgzf_inflate(bgzf.compress(stream[lo:hi], block=block), to_host=False)
        recs, got_tail = pg.bam_index_run(first, 1)
        assert same(recs, exp) and got_tail == tail, (b0, b1, recs.size, exp.size, got_tail, tail)
    stream, table, run, recs, beg, end = _run_setup(pg)
    exp_rows, _, scanned = rpdoracle.discovery(stream, table, 0, beg, end, 500)
    rows, src, irows, isrc, sc = pg.bam_rp_discovery(0, beg, end, 500)
    assert sc == scanned and np.array_equal(rpd_rows(run, recs, rows, src, irows, isrc), exp_rows)
    with pytest.raises(api.PgError):
        pg.bam_index_run(len(run) + 1, 1)


@pytest.mark.parametrize("case", list(G.QUERIES))
def test_matches_reference_dumps(pg, case):
    z = np.load(os.path.join(GOLD, "sr_%s.npz" % case))
    stream = case_stream(case)
    data = open(os.path.join(GOLD, "sim_prefix.bam"), "rb").read() if case == "sim_prefix" else bgzf.compress(stream)
    pg.bgzf_inflate(data, to_host=False)
    recs, _, _ = pg.bam_index()
    for k, (tid, beg, end, isz, min_q) in enumerate(G.QUERIES[case]):
        cands, seqs, refs, _ = pg.bam_sr_candidates(tid, beg, end, isz, int(z["chr_len"][tid]), min_anchor_quality=min_q)
        c, r = rows_of_product(stream, recs, cands, seqs, refs, isz)
        assert np.array_equal(c, z["cand_%d" % k]), ("candidates", case, k, _diff(c, z["cand_%d" % k]))
        assert np.array_equal(r, z["refs_%d" % k]), ("reference-supporting reads", case, k, _diff(r, z["refs_%d" % k]))


def test_matches_oracle_on_a_larger_stream(pg):
    stream = G.paired_stream(30_000, 77)
    table = bam.serial_index(stream)
    pg.bgzf_inflate(bgzf.compress(stream), to_host=False)
    recs, _, _ = pg.bam_index()
    for tid, beg, end, isz, min_q in ((0, 0, 400_000, 500, 0), (1, 50_000, 250_000, 300, 30), (0, 399_000, 400_000, 500, 0)):
        exp_c, exp_r = sroracle.candidates(stream, table, tid, beg, end, isz, G.SYNTH_REFS[tid][1], min_anchor_quality=min_q)
        cands, seqs, refs, scanned = pg.bam_sr_candidates(tid, beg, end, isz, G.SYNTH_REFS[tid][1], min_anchor_quality=min_q)
        c, r = rows_of_product(stream, recs, cands, seqs, refs, isz)
        assert np.array_equal(c, exp_c), ("candidates", tid, beg, end, isz, min_q, _diff(c, exp_c))
        assert np.array_equal(r, exp_r), ("reference-supporting reads", tid, beg, end, isz, min_q, _diff(r, exp_r))
        assert scanned >= cands.size // 3


def test_repeated_names_pair_by_rank(pg):
    for names in ([b"same"] * 7, [b"a", b"ab", b"a", b"abc", b"ab", b"abc", b"a"], [b"r%d" % (i % 5) for i in range(23)]):
        stream = named_stream(names)
        pg.bgzf_inflate(bgzf.compress(stream), to_host=False)
        recs, _, _ = pg.bam_index()
        exp_c, exp_r = sroracle.candidates(stream, bam.serial_index(stream), 0, 0, 100_000, 500, 100_000)
        cands, seqs, refs, scanned = pg.bam_sr_candidates(0, 0, 100_000, 500, 100_000)
        c, r = rows_of_product(stream, recs, cands, seqs, refs, 500)
        assert scanned == len(names), (scanned, names)
        assert np.array_equal(c, exp_c) and np.array_equal(r, exp_r), (names, _diff(c, exp_c), _diff(r, exp_r))


def test_errors(pg):
    fresh = api.PatternGrowth()
    with pytest.raises(api.PgError):
        fresh.bam_sr_candidates(0, 0, 1000, 500, 100_000)              # nothing resident
    stream = named_stream([b"x", b"y", b"x"])
    fresh.bgzf_inflate(bgzf.compress(stream), to_host=False)
    with pytest.raises(api.PgError):
        fresh.bam_sr_candidates(0, 0, 1000, 500, 100_000)              # inflated but not indexed
    fresh.bam_index()
    with pytest.raises(api.PgError):
        fresh.bam_sr_candidates(0, 0, 100_000, 60, 100_000)            # the reference stops the program here
    cands, _, _, _ = fresh.bam_sr_candidates(0, 0, 100_000, 61, 100_000)
    assert cands.size > 0
    fresh.close()


def test_split_read_filter_on_a_run_of_blocks(pg):
    stream, table, run, recs, beg, end = _run_setup(pg)
    exp_c, exp_r = sroracle.candidates(stream, table, 0, beg, end, 500, G.SYNTH_REFS[0][1])
    cands, seqs, refs, _ = pg.bam_sr_candidates(0, beg, end, 500, G.SYNTH_REFS[0][1])
    c, r = rows_of_product(run, recs, cands, seqs, refs, 500)
    assert np.array_equal(c, exp_c) and np.array_equal(r, exp_r), (_diff(c, exp_c), _diff(r, exp_r))


@pytest.mark.skipif(not os.path.exists(DEMO), reason="oracle/_ref/pindel_gpu_demo not built (needs /root/reference at build time)")
def test_reference_host_bam_input_with_the_device_filter(tmp_path):
    env = dict(os.environ, PG_BRIDGE_SR_UNCHAINED="1", LD_LIBRARY_PATH=os.path.join(ROOT, "pindel_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    D.sim_prefix_bam(DEMO, tmp_path, env)


def test_candidates_searched_where_the_filter_left_them(pg):
    """pg_bam_sr_close_end: the filter's rows become the search kernel's batch on the device.  Same rows as pg_bam_sr_candidates,
    and the close ends equal pg_close_end on the same reads uploaded from the host (and therefore the oracle's)."""
    from pindel_b200 import runs as R, synth
    from tests import helpers as H
    stream = G.paired_stream(30_000, 78)
    pg.bgzf_inflate(bgzf.compress(stream), to_host=False)
    recs, _, _ = pg.bam_index()
    refs_np = [synth.make_reference(ln, 900 + i) for i, (_, ln) in enumerate(G.SYNTH_REFS)]
    padded = [synth.pad_reference(r) for r in refs_np]
    pg.set_reference(padded)
    for tid, beg, end, isz in ((0, 0, 400_000, 500), (1, 50_000, 250_000, 300)):
        c0, s0, r0, n0 = pg.bam_sr_candidates(tid, beg, end, isz, G.SYNTH_REFS[tid][1])
        launches = pg.kernel_launches
        cands, seqs, refs, scanned, out, runs = pg.bam_sr_close_end(tid, beg, end, isz, G.SYNTH_REFS[tid][1], chr_id=tid)
        assert pg.kernel_launches == launches + 1           # one search launch for the size-asking call and the fetch together
        assert np.array_equal(cands, c0) and np.array_equal(seqs, s0) and np.array_equal(refs, r0) and scanned == n0
        assert cands.size > 100
        # the same reads through the host-array entry point
        blob = b"".join(bytes(seqs[k, :cands["read_length"][k]]) for k in range(cands.size))
        off = np.zeros(cands.size + 1, dtype=np.uint32)
        off[1:] = np.cumsum(cands["read_length"])
        kw = dict(seq_blob=np.frombuffer(blob, dtype=np.uint8), seq_off=off, matched_d=cands["matched_d"], matched_rel_pos=cands["matched_rel_pos"],
                  insert_size=np.full(cands.size, isz, np.int16), chr_id=np.full(cands.size, tid, np.int32))
        out2, runs2 = pg.get_close_end(**kw)
        a = R.per_read(out, runs, kw["seq_blob"], off)
        b = R.per_read(out2, runs2, kw["seq_blob"], off)
        H.compare(b, a, "filter + search on the device vs pg_close_end from host arrays")
        from oracle import pgoracle, refhook
        st, _ = pgoracle.search(pgoracle.default_params(), padded, kw["seq_blob"], off, kw["matched_d"], kw["matched_rel_pos"], kw["insert_size"],
                                chr_id=kw["chr_id"], far=False, threads=8)
        exp = refhook.parse_stream(st, cands.size)
        for e, g in zip(exp, a):
            assert e["flip"] == g["flip"] and e["L"] == g["L"] and np.array_equal(e["close_raw"], g["close_raw"])
