"""BASELINE.json-sized batches on the GPU, checked through size-independent properties: reads are independent units, so
(1) a random sample of a multi-million-read batch must equal the oracle run on that sample alone, (2) permuting the batch
permutes the results, (3) repeating the launch repeats the records.  configs[1] (100 bp, 10 Mbp, deletions only) and the
configs[2] shape (150 bp, 63 Mbp, D/SI/INV/TD mix) are both exercised; the batches are generated on the device."""
import numpy as np
import pytest

from pindel_b200 import api, runs as R, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def gen(ref_np, n, L, seed, mix):
    import torch
    dev = torch.device("cuda", 0)
    ref_t = torch.from_numpy(ref_np).to(dev)
    seq, md, pos, isz, kind = synth.make_reads_torch(ref_t, n, L, seed, mix=mix)
    return (seq.cpu().numpy(), md.cpu().numpy(), pos.cpu().numpy().astype(np.uint32), isz.cpu().numpy(), kind.cpu().numpy())


def canon(per):
    return [(r["flip"], r["L"], r["close_raw"].tobytes(), r["far"].tobytes()) for r in per]


@pytest.mark.parametrize("name,ref_bases,ref_seed,n,L,mix", [
    ("configs[1]: 1 M x 100 bp, 10 Mbp, deletions only", 10_000_000, 20260925, 1_000_000, 100, (1.0, 0.0, 0.0, 0.0)),
    ("configs[2] shape: 3 M x 150 bp, 63 Mbp, D/SI/INV/TD", 62_435_964, 20260926, 3_000_000, 150, (0.60, 0.25, 0.05, 0.10)),
])
def test_full_size_batch_sample_equals_oracle(name, ref_bases, ref_seed, n, L, mix):
    ref = synth.make_reference(ref_bases, ref_seed)
    padded = synth.pad_reference(ref)
    seq, md, pos, isz, kind = gen(ref, n, L, 4242, mix)
    off = (np.arange(n + 1, dtype=np.uint64) * L).astype(np.uint32)
    pg = api.PatternGrowth()
    pg.set_reference([padded])
    out, runs = pg.search(seq_blob=seq.reshape(-1), seq_off=off, matched_d=md, matched_rel_pos=pos, insert_size=isz)
    assert (out["flags"] & (R.F_TOOLONG | R.F_OVERFLOW)).sum() == 0
    # (3) the same launch again
    out2, runs2 = pg.search(seq_blob=seq.reshape(-1), seq_off=off, matched_d=md, matched_rel_pos=pos, insert_size=isz)
    rng = np.random.default_rng(7)
    sample = np.sort(rng.choice(n, size=20_000, replace=False))
    sub = synth.ReadBatch(seq[sample], md[sample], pos[sample], isz[sample], kind[sample])
    blob_s, off_s = sub.blob()
    got = R.per_read(out[sample], runs, blob_s, off_s)
    got2 = R.per_read(out2[sample], runs2, blob_s, off_s)
    assert canon(got) == canon(got2), name + ": two launches differ"
    # (1) the oracle on the sample alone
    exp = H.oracle_results([padded], sub, threads=8)
    H.compare(exp, got, name)
    # (2) a permuted batch gives permuted results
    perm = rng.permutation(200_000)
    out_p, runs_p = pg.search(seq_blob=seq[perm].reshape(-1), seq_off=off[:perm.size + 1], matched_d=md[perm], matched_rel_pos=pos[perm],
                              insert_size=isz[perm])
    a = canon(R.per_read(out_p[:3000], runs_p, None, None))
    b = canon(R.per_read(out[perm[:3000]], runs, None, None))
    assert a == b, name + ": permutation changed per-read results"
    # sanity of the workload itself: most split reads get both ends
    found = sum(1 for r in got if len(r["far"])) / len(got)
    assert found > 0.6, found
    pg.close()
