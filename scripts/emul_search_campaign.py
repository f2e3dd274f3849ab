"""scripts/emul_search_campaign.py -- CPU-only: the search kernel's warp program (pg_warp.cuh, run by the 32-fiber lane emulator of
tests/emul/) against the oracle port on random option sets, read lengths and references with N runs (--repetitive: with repeats and low-complexity stretches), in the way
tests/test_gpu_fuzz.py does on the hardware.  The emulator runs about 15 k reads per second on 8 host threads; not part of pytest.
    python scripts/emul_search_campaign.py [--rounds N] [--reads M] [--seed S] [--x LO HI] [--ref-len LO HI]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def mixed(a):
    from pindel_b200 import runs as R
    from tests import helpers as H
    from tests.emul import emul
    from tests.test_gpu_fuzz import fuzz_case, oracle_answers, region_array
    lib = emul.build()
    names = dict(seq_error_rate="seq_error_rate", sensitivity="sensitivity", max_mismatch_rate="max_mismatch_rate",
                 additional_mismatch="add_mm", min_close="min_close", min_perfect_match="min_perfect", max_range_index="max_range_index")
    total = 0
    for r in range(a.rounds):
        seed = a.seed + r
        settings, padded, blob, off, md, pos, isz, chr_id, regions, roff, lens = fuzz_case(seed, n_per=a.reads)
        prm = emul.make_params(lib, **{names[k]: v for k, v in settings.items()})
        out, pool = emul.search(lib, prm, padded, blob, off, md, pos, isz, chr_id=chr_id, regions=region_array(regions), region_off=roff)
        assert (out["flags"] & (R.F_TOOLONG | R.F_OVERFLOW)).sum() == 0
        H.compare(oracle_answers(settings, padded, blob, off, md, pos, isz, chr_id, regions, roff), R.per_read(out, pool, blob, off),
                  "mixed seed %d %s lens %s" % (seed, settings, lens.tolist()))
        total += len(off) - 1
        print("round %d ok: seed %d, %d chromosomes, lens %s, %d clusters, %s" % (r, seed, len(padded), lens.tolist(), len(regions), settings), flush=True)
    print("campaign clean: %d rounds, %d reads, mixed lengths / chromosomes / clusters, warp program (emulated) == oracle port" % (a.rounds, total))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=12)
    ap.add_argument("--reads", type=int, default=400)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--x", type=int, nargs=2, default=[0, 4], help="range of -x (max_range_index), inclusive")
    ap.add_argument("--ref-len", type=int, nargs=2, default=[250_000, 400_000])
    ap.add_argument("--repetitive", action="store_true", help="references with near-identical repeats, low-complexity stretches and N runs "
                    "(tests/test_emulator_filter.py), reads with N: what stresses the far rings' filter")
    ap.add_argument("--mixed", action="store_true", help="the GPU fuzz sweep's cases instead (tests/test_gpu_fuzz.py: mixed read lengths "
                    "40..300 in one batch, 1-3 chromosomes, random clusters also on other chromosomes); --reads = reads per length and chromosome")
    a = ap.parse_args()
    if a.mixed:
        return mixed(a)
    from oracle import pgoracle, refhook
    from pindel_b200 import runs as R, synth
    from tests import helpers as H
    from tests.emul import emul
    lib = emul.build()
    rng = np.random.default_rng(a.seed)
    total = 0
    for r in range(a.rounds):
        st = dict(max_range_index=int(rng.integers(a.x[0], a.x[1] + 1)), add_mm=int(rng.integers(1, 4)), min_perfect=int(rng.integers(0, 8)),
                  min_close=int(rng.integers(4, 14)), max_mismatch_rate=float(rng.choice([0.0, 0.01, 0.02, 0.05, 0.1])),
                  seq_error_rate=float(rng.choice([0.005, 0.01, 0.02, 0.04])), sensitivity=float(rng.choice([0.9, 0.95, 0.99])))
        L = int(rng.choice([40, 64, 75, 100, 101, 150, 200, 250]))
        if a.repetitive:
            from tests.test_emulator_filter import repetitive_reference
            ref = repetitive_reference(int(rng.integers(a.ref_len[0], a.ref_len[1])), int(rng.integers(1, 1 << 30)))
            batch = synth.make_reads(ref, a.reads, L, seed=int(rng.integers(1, 1 << 30)), max_del=int(rng.choice([1000, 6000, 20000])), n_rate=0.004)
        else:
            ref = synth.make_reference(int(rng.integers(a.ref_len[0], a.ref_len[1])), int(rng.integers(1, 1 << 30)),
                                       n_runs=[(int(rng.integers(20_000, 200_000)), int(rng.integers(1, 3000)))])
            batch = synth.make_reads(ref, a.reads, L, seed=int(rng.integers(1, 1 << 30)), max_del=int(rng.choice([200, 1000, 20000])))
        padded = synth.pad_reference(ref)
        blob, off = batch.blob()
        prm = emul.make_params(lib, **st)
        out, pool = emul.search(lib, prm, [padded], blob, off, batch.matched_d, batch.matched_rel_pos, batch.insert_size)
        assert (out["flags"] & (R.F_TOOLONG | R.F_OVERFLOW)).sum() == 0
        got = R.per_read(out, pool, blob, off)
        P = pgoracle.default_params(max_range_index=st["max_range_index"], additional_mismatch=st["add_mm"], min_perfect_match=st["min_perfect"],
                                    min_close=st["min_close"], max_mismatch_rate=st["max_mismatch_rate"], seq_error_rate=st["seq_error_rate"],
                                    sensitivity=st["sensitivity"])
        stream, _ = pgoracle.search(P, [padded], blob, off, batch.matched_d, batch.matched_rel_pos, batch.insert_size, threads=8)
        H.compare(refhook.parse_stream(stream, batch.n), got, "round %d L=%d %s" % (r, L, st))
        total += batch.n
        print("round %d ok: L=%d %s" % (r, L, st), flush=True)
    print("campaign clean: %d rounds, %d reads, warp program (emulated) == oracle port" % (a.rounds, total))


if __name__ == "__main__":
    main()
