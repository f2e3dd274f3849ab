"""Error behaviour of the C ABI on a device: bad arguments and unsupported inputs come back as error codes with a message
(the host keeps Pindel's print-and-exit convention), never as silent partial results."""
import numpy as np
import pytest

from pindel_b200 import api, runs as R, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small():
    ref = synth.make_reference(300_000, 81)
    return ref, synth.pad_reference(ref), synth.make_reads(ref, 200, 100, seed=82)


def test_search_before_reference_is_a_state_error(small):
    ref, padded, batch = small
    pg = api.PatternGrowth()
    with pytest.raises(api.PgError, match="no reference"):
        pg.search(**H.batch_kwargs(batch))
    pg.close()


def test_missing_anchor_fields_and_close_info(small):
    ref, padded, batch = small
    kw = H.batch_kwargs(batch)
    pg = api.PatternGrowth()
    pg.set_reference([padded])
    with pytest.raises(api.PgError, match="needs matched_d"):
        pg.search(seq_blob=kw["seq_blob"], seq_off=kw["seq_off"])
    with pytest.raises(api.PgError, match="close_info"):
        pg.search_far_ends(seq_blob=kw["seq_blob"], seq_off=kw["seq_off"])
    pg.close()


def test_out_of_range_chromosome_and_long_reads(small):
    ref, padded, batch = small
    kw = H.batch_kwargs(batch)
    pg = api.PatternGrowth()
    pg.set_reference([padded])
    with pytest.raises(api.PgError, match="chr_id out of range"):
        pg.search(chr_id=np.full(batch.n, 3, np.int32), **kw)
    long_seq = np.frombuffer(b"ACGT" * 150, dtype=np.uint8)          # 600 bases: g_maxMismatch has 500 entries
    with pytest.raises(api.PgError, match="499"):
        pg.search(seq_blob=long_seq, seq_off=np.array([0, 600], np.uint32), matched_d=np.array([ord("+")], np.uint8),
                  matched_rel_pos=np.array([5000], np.uint32), insert_size=np.array([500], np.int16))
    # the context stays usable after an error
    out, runs = pg.search(**kw)
    H.compare(H.oracle_results([padded], batch, threads=2), R.per_read(out, runs, kw["seq_blob"], kw["seq_off"]), "after errors")
    pg.close()


def test_bad_settings_are_rejected():
    with pytest.raises(api.PgError, match="out of range"):
        api.PatternGrowth(max_range_index=12)
    with pytest.raises(api.PgError, match="device ordinal"):
        api.PatternGrowth(device=99)


def test_too_many_levels_is_unsupported(small):
    ref, padded, batch = small
    pg = api.PatternGrowth(seq_error_rate=0.2, additional_mismatch=20)       # > 32 mismatch levels for 100 bp reads
    pg.set_reference([padded])
    with pytest.raises(api.PgError, match="32 mismatch levels"):
        pg.search(**H.batch_kwargs(batch))
    pg.close()


def test_bad_read_in_a_later_chunk_of_a_chunked_call(small, monkeypatch):
    """The chunked path checks every chunk's reads right before their upload (behind the work of the chunks before): an
    out-of-range chromosome in the last chunk is reported after earlier chunks were launched, and the context serves the next
    call."""
    ref, padded, batch = small
    monkeypatch.setenv("PG_PIPE_CHUNK", "256")
    big = synth.make_reads(ref, 2000, 100, seed=83)
    kw = H.batch_kwargs(big)
    pg = api.PatternGrowth()
    pg.set_reference([padded])
    good = R.per_read(*pg.search(**kw), kw["seq_blob"], kw["seq_off"])
    chr_id = np.zeros(big.n, dtype=np.int32)
    chr_id[1900] = 5
    with pytest.raises(api.PgError, match="chr_id out of range"):
        pg.search(chr_id=chr_id, **kw)
    again = R.per_read(*pg.search(**kw), kw["seq_blob"], kw["seq_off"])
    H.compare(good, again, "after a failed chunked call")
    # mixed read lengths: every chunk picks its kernel variant from its own longest read
    long_batch = synth.make_reads(ref, 600, 300, seed=84)
    blob = np.concatenate([kw["seq_blob"], long_batch.seq.reshape(-1)])
    off = np.concatenate([kw["seq_off"], kw["seq_off"][-1] + np.arange(1, long_batch.n + 1, dtype=np.uint32) * 300]).astype(np.uint32)
    mixed = dict(seq_blob=blob, seq_off=off, matched_d=np.concatenate([big.matched_d, long_batch.matched_d]),
                 matched_rel_pos=np.concatenate([big.matched_rel_pos, long_batch.matched_rel_pos]),
                 insert_size=np.concatenate([big.insert_size, long_batch.insert_size]))
    a = R.per_read(*pg.search(**mixed), blob, off)
    pg.close()
    monkeypatch.setenv("PG_PIPE_CHUNK", "1000000")
    pg1 = api.PatternGrowth()
    pg1.set_reference([padded])
    b = R.per_read(*pg1.search(**mixed), blob, off)
    pg1.close()
    H.compare(b, a, "chunk-wise kernel variants vs one launch")
