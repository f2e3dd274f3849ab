"""GPU parity of the SV-classification stage: pg_classify through the C ABI against the
reference's own dumps (tests/golden/sv_*.npz), against the oracle on fresh seeded batches, and chained behind the GPU search."""
import numpy as np
import pytest

from oracle import pgoracle, svoracle
from pindel_b200 import api, runs as R, sv, synth
from tests import helpers as H, sv_helpers as S
from tests.golden.make_sv_golden import CASES

pytestmark = pytest.mark.gpu


def settings_of(c):
    return dict(c["case"].settings)


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("raw_close", [False, True])
def test_classify_matches_reference_dumps(name, raw_close):
    c = S.load_sv_case(name)
    pg = api.PatternGrowth(**settings_of(c))
    pg.set_reference([c["case"].padded])
    b = S.product_batch(c, raw_close=raw_close)
    rec = pg.classify(b, c["win"])
    assert pg.classify_launches >= 1
    got = S.dump_from_records(rec, c["seqs"], b["rflags"], driver_mask=c["win"]["driver_mask"])
    svoracle.compare_dumps(c["golden"], got, name)
    pg.close()


@pytest.mark.parametrize("L,n,seed,region_frac,wend_frac", [(100, 30000, 31, (0.0, 1.0), 0.7), (150, 20000, 32, (0.2, 0.5), 1.0),
                                                            (60, 20000, 33, (0.0, 1.0), 1.0)])
def test_search_then_classify_matches_oracle(L, n, seed, region_frac, wend_frac):
    """GPU search -> GPU classification, against oracle search -> oracle drivers, on a fresh seeded batch."""
    ref = synth.make_reference(3_000_000, 777 + seed, n_runs=[(1_000_000, 3000)])
    padded = synth.pad_reference(ref)
    batch = synth.make_reads(ref, n, L, seed=seed)
    pg = api.PatternGrowth()
    pg.set_reference([padded])
    bk = H.batch_kwargs(batch)
    out, runs = pg.search(**bk)
    got_search = R.per_read(out, runs, bk["seq_blob"], bk["seq_off"])
    exp_search = H.oracle_results([padded], batch)
    H.compare(exp_search, got_search, "search")
    seqs = S.final_seqs(bk["seq_blob"], bk["seq_off"], got_search)
    ci = S.classify_inputs(exp_search, seqs, batch.matched_d, batch.insert_size)
    idx = ci["index"]
    region = (int(ref.size * region_frac[0]), int(ref.size * region_frac[1]))
    win = sv.window(int(padded.size), int(ref.size * wend_frac), region=region)
    P = svoracle.make_params(pgoracle.prob_table(), win["window_end"], win["num_boxes"], win["box_size"],
                             region=((1, region[0]), (1, region[1])))
    kw = {k: ci[k] for k in ("seq_blob", "seq_off", "matched_d", "insert_size", "rflags", "up_rows", "up_off", "n_close")}
    exp = svoracle.classify(P, padded.tobytes(), **kw)
    # the product side takes the search kernel's own runs: no expansion, raw close lists (cleaned on the device)
    sq = [seqs[i] for i in idx]
    run_off = np.zeros(len(idx) + 1, dtype=np.uint32)
    parts = []
    for k, i in enumerate(idx):
        o = out[i]
        a, cnt = int(o["run_off"]), int(o["n_close_runs"]) + int(o["n_far_runs"])
        parts.append(runs[a:a + cnt])
        run_off[k + 1] = run_off[k] + cnt
    b = dict(seq_blob=ci["seq_blob"], seq_off=ci["seq_off"], matched_d=ci["matched_d"], insert_size=ci["insert_size"],
             rflags=ci["rflags"], run_off=run_off, n_close_runs=out["n_close_runs"][idx].astype(np.uint16),
             runs=np.concatenate(parts))
    rec = pg.classify(b, win)
    got = S.dump_from_records(rec, sq, b["rflags"])
    svoracle.compare_dumps(exp, got, "chained L=%d" % L)

    def remap(rec_all):
        """records of a whole-batch classification -> indices among the reads with a close end (the oracle's batch)"""
        pos = np.searchsorted(idx, rec_all["read"])
        assert np.all(pos < len(idx)) and np.array_equal(idx[pos], rec_all["read"]), "record for a read without a close end"
        r2 = rec_all.copy()
        r2["read"] = pos
        return r2

    # (a) resident after the close+far launch: nothing uploaded again, flips applied on the device
    out2, runs2 = pg.search(**bk)
    assert np.array_equal(out2, out) or np.array_equal(out2["n_far_runs"], out["n_far_runs"])
    rec_a = pg.classify_resident(win)
    svoracle.compare_dumps(exp, S.dump_from_records(remap(rec_a), sq, b["rflags"]), "resident after pg_search L=%d" % L)
    # (b) the reference's call order: close ends first, then SearchFarEnds on the reads that have one, then the drivers
    ci_arr = np.zeros(len(idx), dtype=R.CLOSEINFO_DT)
    close_off = np.zeros(len(idx) + 1, dtype=np.uint32)
    cparts = []
    for k, i in enumerate(idx):
        c = exp_search[i]["close"]
        ci_arr[k] = (int(c[-1, 1]) & 0xFFFFFFFF, int(c[-1, 0]), 0)
        o = out[i]
        cparts.append(runs[int(o["run_off"]):int(o["run_off"]) + int(o["n_close_runs"])])
        close_off[k + 1] = close_off[k] + int(o["n_close_runs"])
    pg.search_far_ends(seq_blob=ci["seq_blob"], seq_off=ci["seq_off"], matched_d=ci["matched_d"],
                       matched_rel_pos=np.asarray(batch.matched_rel_pos)[idx], insert_size=ci["insert_size"], close_info=ci_arr)
    rec_b = pg.classify_resident(win, rflags=b["rflags"], close_run_off=close_off, close_runs=np.concatenate(cparts))
    svoracle.compare_dumps(exp, S.dump_from_records(rec_b, sq, b["rflags"]), "resident after pg_far_end L=%d" % L)
    pg.close()


def test_classify_resident_needs_a_resident_far_end_launch():
    pg = api.PatternGrowth()
    c = S.load_sv_case("synth_L250_default")
    pg.set_reference([c["case"].padded])
    with pytest.raises(api.PgError):
        pg.classify_resident(c["win"])          # nothing launched
    ref = synth.make_reference(300_000, 5)
    batch = synth.make_reads(ref, 500, 100, seed=6)
    pg.set_reference([synth.pad_reference(ref)])
    pg.get_close_end(**H.batch_kwargs(batch))
    with pytest.raises(api.PgError):
        pg.classify_resident(sv.window(500_000, 300_000))   # a close-end-only launch has no far ends
    pg.close()


def test_classify_argument_errors():
    pg = api.PatternGrowth()
    c = S.load_sv_case("synth_L250_default")
    b = S.product_batch(c)
    with pytest.raises(api.PgError):
        pg.classify(b, c["win"])            # no reference yet
    pg.set_reference([c["case"].padded])
    bad = dict(c["win"], box_size=0)
    with pytest.raises(api.PgError):
        pg.classify(b, bad)
    bad = dict(c["win"], chr=3)
    with pytest.raises(api.PgError):
        pg.classify(b, bad)
    small = np.empty(3, dtype=sv.RECORD_DT)   # too small: the wrapper repeats the call with the size that came back
    rec = pg.classify(b, c["win"], records=small)
    assert rec.size > 3
    pg.close()
