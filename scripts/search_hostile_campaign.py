"""scripts/search_hostile_campaign.py -- CPU-only, meant to run under AddressSanitizer (PG_EMUL_CXXFLAGS, see the end of this text):
the search kernel's warp program (pg_warp.cuh on the lane emulator) on batches whose per-read fields are HOSTILE -- anchor
positions anywhere in 32 bits, insert sizes negative / zero / 32767, anchor strands that are neither '+' nor '-', reads of length
0..300 made of arbitrary bytes, clusters with start > end or far outside their chromosome or on the wrong chromosome, far-end-only
batches with arbitrary close ends.  The reference has no defined behaviour for such input (it indexes outside its strings); the
library's contract is that every window is clipped to its chromosome and nothing outside the packed reference, the batch or the
run pool is touched.  What is checked: no sanitizer report, no crash, termination, run offsets inside the pool.
    PG_EMUL_CXXFLAGS="-fsanitize=address,undefined -fno-omit-frame-pointer" LD_PRELOAD="$(g++ -print-file-name=libasan.so) $(g++ -print-file-name=libubsan.so)" \\
    ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=halt_on_error=1 python scripts/search_hostile_campaign.py [--rounds N]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=200)
    ap.add_argument("--reads", type=int, default=300)
    ap.add_argument("--seed", type=int, default=1)
    run(ap.parse_args())


def run(a):
    from pindel_b200 import runs as R, synth
    from tests.emul import emul
    lib = emul.build(force=bool(os.environ.get("PG_EMUL_CXXFLAGS")))
    rng = np.random.default_rng(a.seed)
    total = flagged = with_runs = 0
    for r in range(a.rounds):
        n_chr = int(rng.integers(1, 4))
        refs = [synth.make_reference(int(rng.choice([300, 5000, 250_000])), int(rng.integers(1, 1 << 30))) for _ in range(n_chr)]
        padded = [synth.pad_reference(x) for x in refs]
        n = a.reads
        lens = rng.choice([0, 1, 7, 8, 31, 32, 33, 100, 150, 256, 257, 300], size=n)
        off = np.zeros(n + 1, dtype=np.uint32)
        off[1:] = np.cumsum(lens)
        kind = int(rng.integers(0, 3))
        alphabet = [np.frombuffer(b"ACGTN", dtype=np.uint8), np.arange(256, dtype=np.uint8), np.frombuffer(b"ACGTNacgtn=MRSVWYHKDB\0 ", dtype=np.uint8)][kind]
        blob = rng.choice(alphabet, size=int(off[-1])).astype(np.uint8)
        md = rng.choice(np.frombuffer(b"+-+-+-?\0", dtype=np.uint8), size=n)
        plen = np.array([p.size for p in padded], dtype=np.int64)
        chr_id = rng.integers(0, n_chr, n).astype(np.int32)
        pos = np.where(rng.random(n) < 0.5, rng.integers(0, plen[chr_id] + 1000), rng.integers(0, 1 << 32, n, dtype=np.uint64)).astype(np.uint32)
        isz = rng.choice(np.array([-32768, -500, -1, 0, 1, 200, 500, 32767], dtype=np.int16), size=n)
        if refs[0].size >= 250_000:                       # half of the batch: real reads of chromosome 0, every field damaged on a part of them
            L = int(rng.choice([40, 100, 150, 250]))
            real = synth.make_reads(refs[0], n // 2, L, seed=int(rng.integers(1, 1 << 30)), max_del=3000)
            seqs = [real.seq[i].tobytes() for i in range(real.n)] + [blob[off[i]:off[i + 1]].tobytes() for i in range(n // 2, n)]
            off = np.zeros(n + 1, dtype=np.uint32)
            off[1:] = np.cumsum([len(x) for x in seqs])
            blob = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy()
            keep = rng.random(n // 2) < 0.6
            h = n // 2
            md[:h] = np.where(keep, real.matched_d, md[:h])
            pos[:h] = np.where(keep, real.matched_rel_pos, pos[:h])
            isz[:h] = np.where(rng.random(h) < 0.6, real.insert_size, isz[:h])
            chr_id[:h] = np.where(rng.random(h) < 0.8, 0, chr_id[:h])
            hit = rng.random(blob.size) < 0.002
            blob[hit] = rng.integers(0, 256, int(hit.sum()), dtype=np.uint8)
        regions, roff = [], [0]
        for i in range(n):
            for _ in range(int(rng.choice([0, 0, 1, 3]))):
                c = int(rng.integers(0, n_chr))
                s = int(rng.choice([0, 1, int(rng.integers(0, plen[c] + 5000)), 0xFFFFFF00, 0x7FFFFFFF]))
                e = int(rng.choice([s, s + 1, s + int(rng.integers(0, 3000)), max(0, s - 50), 0xFFFFFFFF]))
                regions.append((c, s, e))
            roff.append(len(regions))
        rg = np.zeros(len(regions), dtype=R.REGION_DT)
        if regions:
            arr = np.array(regions, dtype=np.uint64)
            rg["chr"], rg["start"], rg["end"] = arr[:, 0].astype(np.int32), arr[:, 1].astype(np.uint32), arr[:, 2].astype(np.uint32)
        roff = np.array(roff, dtype=np.uint32)
        mode = int(rng.choice([3, 3, 1, 2]))
        ci = None
        if mode == 2:
            ci = np.zeros(n, dtype=R.CLOSEINFO_DT)
            ci["last_abs_loc"] = np.where(rng.random(n) < 0.5, rng.integers(0, plen[chr_id] + 1000), rng.integers(0, 1 << 32, n, dtype=np.uint64)).astype(np.uint32)
            ci["max_len"] = rng.choice([0, 1, 8, 50, 100, 300, 65535], size=n)
        st = dict(max_range_index=int(rng.integers(0, 5)), add_mm=int(rng.integers(1, 4)), min_perfect=int(rng.integers(0, 8)),
                  min_close=int(rng.integers(1, 14)), max_mismatch_rate=float(rng.choice([0.0, 0.02, 0.1, 1.0])))
        prm = emul.make_params(lib, **st)
        out, pool = emul.search(lib, prm, padded, blob, off, md, pos, isz, chr_id=chr_id, mode=mode, close_info=ci, regions=rg, region_off=roff,
                                pool_per_read=700)
        ok = (out["flags"] & (R.F_TOOLONG | R.F_OVERFLOW)) == 0
        assert np.all(out["run_off"][ok].astype(np.int64) + out["n_close_runs"][ok].astype(np.int64) + out["n_far_runs"][ok].astype(np.int64) <= pool.size)
        total += n
        flagged += int((~ok).sum())
        with_runs += int(((out["n_close_runs"] > 0) & ok).sum())
        if r % 20 == 0:
            print("round %d ok" % r, flush=True)
    print("campaign clean: %d rounds, %d hostile reads (%d came back with runs, %d refused by flag); no crash%s" %
          (a.rounds, total, with_runs, flagged, ", no sanitizer report" if "asan" in os.environ.get("LD_PRELOAD", "") else ""))
    return total, with_runs


if __name__ == "__main__":
    main()
