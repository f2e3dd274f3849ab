"""GPU parity tests proper: the CUDA path, called through the C ABI (pindel_b200.api -> libpindel_pg.so), against
the oracle on the same seeded inputs.  Bit-exact: positions, lengths, directions, strands, mismatch counts,
flip state, read length after the search."""
import numpy as np
import pytest

from pindel_b200 import api, runs as R, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref2m():
    ref = synth.make_reference(2_000_000, 20260925, n_runs=[(700_000, 5000), (1_200_000, 37), (1_500_003, 1)])
    return ref, synth.pad_reference(ref)


def run_gpu(pg, batch, **kw):
    bk = H.batch_kwargs(batch)
    out, runs = pg.search(**bk, **kw)
    return R.per_read(out, runs, bk["seq_blob"], bk["seq_off"]), out, runs


@pytest.mark.parametrize("L,n,seed", [(100, 20000, 11), (150, 20000, 12), (75, 10000, 13), (250, 4000, 14)])
def test_fused_search_matches_oracle(ref2m, L, n, seed):
    ref, padded = ref2m
    batch = synth.make_reads(ref, n, L, seed=seed)
    pg = api.PatternGrowth()
    pg.set_reference([padded])
    got, out, runs = run_gpu(pg, batch)
    exp = H.oracle_results([padded], batch)
    H.compare(exp, got, "fused L=%d" % L)
    assert pg.kernel_launches >= 1
    assert (out["flags"] & (R.F_TOOLONG | R.F_OVERFLOW)).sum() == 0
    pg.close()


def test_close_then_far_equals_fused(ref2m):
    ref, padded = ref2m
    batch = synth.make_reads(ref, 8000, 100, seed=21)
    pg = api.PatternGrowth()
    pg.set_reference([padded])
    fused, _, _ = run_gpu(pg, batch)
    kw = H.batch_kwargs(batch)
    out_c, runs_c = pg.get_close_end(**kw)
    close = R.per_read(out_c, runs_c)
    # host side of the seam: flip the sequences that GetCloseEnd flipped, derive close_info
    seqs = batch.seq.copy()
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    ci = np.zeros(batch.n, dtype=R.CLOSEINFO_DT)
    for i, r in enumerate(close):
        if r["flip"]:
            seqs[i] = comp[seqs[i][::-1]]
        if len(r["close"]):
            ci["last_abs_loc"][i] = r["close"][-1][1]
            ci["max_len"][i] = r["close"][-1][0]
    out_f, runs_f = pg.search_far_ends(seq_blob=seqs.reshape(-1), seq_off=kw["seq_off"], close_info=ci)
    far = R.per_read(out_f, runs_f)
    for i in range(batch.n):
        assert np.array_equal(close[i]["close_raw"], fused[i]["close_raw"]), i
        assert np.array_equal(far[i]["far"], fused[i]["far"]), i
    pg.close()


@pytest.mark.parametrize("settings", [dict(max_range_index=4, max_mismatch_rate=0.05),
                                      dict(additional_mismatch=2, min_perfect_match=5, min_close=10, seq_error_rate=0.03)])
def test_non_default_settings(ref2m, settings):
    ref, padded = ref2m
    batch = synth.make_reads(ref, 6000, 100, seed=31, max_del=20000)
    pg = api.PatternGrowth(**settings)
    pg.set_reference([padded])
    got, out, _ = run_gpu(pg, batch)
    exp = H.oracle_results([padded], batch, **settings)
    H.compare(exp, got, str(settings))
    pg.close()


def test_table_matches_oracle():
    from oracle import pgoracle
    pg = api.PatternGrowth()
    assert np.array_equal(pg.max_mismatch_table(), pgoracle.prob_table())
    pg.close()


def test_pipelined_large_batch_path_equals_single_launch(ref2m, monkeypatch):
    """Batches of >= 2 chunks go through the chunked, copy/compute-overlapped path of pg_search; with a small chunk size
    (PG_PIPE_CHUNK) the same code runs on a test-sized batch and must return exactly what a single launch returns."""
    ref, padded = ref2m
    batch = synth.make_reads(ref, 7003, 100, seed=41)
    kw = H.batch_kwargs(batch)
    pg1 = api.PatternGrowth()
    pg1.set_reference([padded])
    a = R.per_read(*pg1.search(**kw), kw["seq_blob"], kw["seq_off"])
    launches_single = pg1.kernel_launches
    stats_single = pg1.last_stats()
    pg1.close()
    monkeypatch.setenv("PG_PIPE_CHUNK", "1000")
    pg2 = api.PatternGrowth()
    pg2.set_reference([padded])
    out, runs = pg2.search(**kw)
    b = R.per_read(out, runs, kw["seq_blob"], kw["seq_off"])
    assert pg2.kernel_launches >= 8 and launches_single == 1
    H.compare(a, b, "pipelined vs single launch")
    # a run buffer that is too small is reported with the required size, then the repeated call succeeds
    small = np.empty(10, dtype=R.RUN_DT)
    out2, runs2 = pg2.search(runs=small, **kw)
    H.compare(a, R.per_read(out2, runs2, kw["seq_blob"], kw["seq_off"]), "pipelined after RUNS_CAP retry")
    # a chunked call leaves batch and results resident exactly like upload + launch: same totals, same records
    assert pg2.last_stats() == stats_single
    assert pg2.device_results()[2] == runs2.size
    out3, runs3 = pg2.download()
    H.compare(a, R.per_read(out3, runs3, kw["seq_blob"], kw["seq_off"]), "download after a chunked call")
    # the same through the call that copies nothing back, and the result slots alternate
    pg2.search_resident(**kw)
    out4, runs4 = pg2.download()
    H.compare(a, R.per_read(out4, runs4, kw["seq_blob"], kw["seq_off"]), "search_resident + download")
    assert pg2.device_results_previous()[2] == runs2.size
    pg2.close()


def test_device_group_equals_single_device(ref2m, monkeypatch):
    """pg_group: the reads of a batch dealt to several contexts in contiguous ranges, records written to their final place
    with global run offsets.  Three contexts on device 0 exercise the splitting and merging on a one-GPU box (with two or
    more GPUs tests/test_gpu_multi.py runs real devices); chunked uploads inside (PG_PIPE_CHUNK) and a too-small run buffer."""
    import torch
    ref, padded = ref2m
    batch = synth.make_reads(ref, 9001, 100, seed=43)
    kw = H.batch_kwargs(batch)
    pg1 = api.PatternGrowth()
    pg1.set_reference([padded])
    a = R.per_read(*pg1.search(**kw), kw["seq_blob"], kw["seq_off"])
    ac = R.per_read(*pg1.get_close_end(**kw), kw["seq_blob"], kw["seq_off"])
    pg1.close()
    monkeypatch.setenv("PG_PIPE_CHUNK", "1000")
    nd = torch.cuda.device_count()
    for devices in ([0, 0, 0], list(range(nd)) if nd > 1 else [0]):
        g = api.PatternGrowthGroup(devices)
        g.set_reference([padded])
        H.compare(a, R.per_read(*g.search(**kw), kw["seq_blob"], kw["seq_off"]), "group %s search" % devices)
        H.compare(ac, R.per_read(*g.get_close_end(**kw), kw["seq_blob"], kw["seq_off"]), "group %s close end" % devices)
        small = np.empty(10, dtype=R.RUN_DT)
        H.compare(a, R.per_read(*g.search(runs=small, **kw), kw["seq_blob"], kw["seq_off"]), "group %s after RUNS_CAP" % devices)
        g.close()
