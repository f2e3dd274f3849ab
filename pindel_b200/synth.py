"""Synthetic references and anchored split-read batches (SURVEY.md section 8(d) read model).

Everything here is deterministic given the seed (numpy PCG64), vectorised, and free of any
dependency on /root/reference, so the same inputs can be regenerated on the GPU box, fed to
the CUDA path, to the C oracle and (where it was built) to the reference hook library.

Read model: a split read spans one planted breakpoint at an offset k in [20, L-20]; the anchor
mate sits at the expected insert distance on either strand; `matched_rel_pos` follows the
reference's build_record_SR convention (src/reader.cpp:857-866: leftmost position for '+'
anchors, position + aligned length for '-' anchors); 1 % substitution errors, 0.1 % N,
10 % non-split decoys that map end to end (they must be rejected by CheckMismatches,
src/searcher.cpp:331-388).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SPACER = 100000  # g_SpacerBeforeAfter, src/pindel.h:122
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b
_CODE = np.zeros(256, dtype=np.uint8)
_CODE[ord("C")], _CODE[ord("G")], _CODE[ord("T")] = 1, 2, 3


def make_reference(n_bases: int, seed: int, n_runs=()) -> np.ndarray:
    """Uniform i.i.d. ACGT reference as a uint8 ASCII array (biological coordinates, unpadded).

    n_runs: iterable of (start, length) stretches overwritten with 'N'."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ref = _ACGT[rng.integers(0, 4, size=n_bases, dtype=np.uint8)]
    for s, ln in n_runs:
        ref[s:s + ln] = ord("N")
    return ref


def pad_reference(ref: np.ndarray) -> np.ndarray:
    """The reference's in-memory chromosome: SPACER 'N' on both sides (src/pindel.cpp:305-309)."""
    out = np.full(ref.size + 2 * SPACER, ord("N"), dtype=np.uint8)
    out[SPACER:SPACER + ref.size] = ref
    return out


@dataclass
class ReadBatch:
    """SoA batch of anchored unmapped read-ends (the inputs of GetCloseEnd, src/pindel.h:523)."""
    seq: np.ndarray            # uint8 [n, L] ASCII (UnmatchedSeq)
    matched_d: np.ndarray      # uint8 [n]  '+' or '-'      (SPLIT_READ::MatchedD)
    matched_rel_pos: np.ndarray  # uint32 [n]                (SPLIT_READ::MatchedRelPos)
    insert_size: np.ndarray    # int16 [n]                   (SPLIT_READ::InsertSize)
    kind: np.ndarray           # uint8 [n] 0=D 1=SI 2=INV 3=TD 4=decoy (generator's truth, not an input)

    @property
    def n(self) -> int:
        return int(self.seq.shape[0])

    @property
    def read_len(self) -> int:
        return int(self.seq.shape[1])

    def slice(self, lo: int, hi: int) -> "ReadBatch":
        return ReadBatch(self.seq[lo:hi], self.matched_d[lo:hi], self.matched_rel_pos[lo:hi],
                         self.insert_size[lo:hi], self.kind[lo:hi])

    def blob(self):
        """(seq_blob bytes-like uint8 [n*L], seq_off uint32 [n+1]) as the C-ABI takes them."""
        n, L = self.seq.shape
        off = (np.arange(n + 1, dtype=np.uint64) * L).astype(np.uint32)
        return np.ascontiguousarray(self.seq).reshape(-1), off

    def to_pindel_text(self, chr_name: str, tag: str = "S") -> str:
        """Pindel text records (src/pindel_read_reader.cpp:50-66)."""
        lines = []
        for i in range(self.n):
            lines.append("@r%d" % i)
            lines.append(self.seq[i].tobytes().decode())
            lines.append("%s\t%s\t%d\t60\t%d\t%s" % (chr(self.matched_d[i]), chr_name,
                                                     int(self.matched_rel_pos[i]), int(self.insert_size[i]), tag))
        return "\n".join(lines) + "\n"


def make_reads(ref: np.ndarray, n: int, read_len: int, seed: int, insert_size: int = 500,
               mix=(0.60, 0.25, 0.05, 0.10), decoy_frac: float = 0.10, max_del: int = 1000,
               sub_rate: float = 0.01, n_rate: float = 0.001, chunk: int = 1 << 18,
               lo: int | None = None, hi: int | None = None) -> ReadBatch:
    """n anchored split reads over `ref` (unpadded ASCII array); events planted in [lo, hi).

    mix = fractions of (deletion, short insertion 1-50 bp, inversion >= 50 bp, tandem dup)
    among the non-decoy reads; deletion sizes are log-uniform in [1, max_del]."""
    L = read_len
    G = ref.size
    margin = 4 * insert_size + 3 * L + max_del + 4096
    lo = margin if lo is None else max(lo, margin)
    hi = G - margin if hi is None else min(hi, G - margin)
    if hi <= lo:
        raise ValueError("reference too small for the requested read model")
    rng = np.random.Generator(np.random.PCG64(seed))
    seqs = np.empty((n, L), dtype=np.uint8)
    md = np.empty(n, dtype=np.uint8)
    mpos = np.empty(n, dtype=np.uint32)
    kinds = np.empty(n, dtype=np.uint8)
    cum = np.cumsum(np.asarray(mix, dtype=np.float64))
    cum /= cum[-1]
    ar = np.arange(L, dtype=np.int64)[None, :]
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        u = rng.random(m)
        kind = np.searchsorted(cum, rng.random(m), side="right").astype(np.uint8)
        kind[u < decoy_frac] = 4
        k = rng.integers(20, L - 20 + 1, size=m)              # bases before the breakpoint
        a = rng.integers(lo, hi, size=m)                       # breakpoint: first segment is ref[a-k, a)
        dsz = np.exp(rng.random(m) * np.log(max_del)).astype(np.int64)  # log-uniform 1..max_del
        isz = rng.integers(1, 51, size=m)                      # short insertion length
        vsz = rng.integers(50, max_del + 1, size=m)            # inversion / tandem-dup span
        # Model geometry ('+' anchor: close end = first segment, read forward).  For '-' anchors the
        # same model is laid out on the reverse-complement strand (coordinate x -> G-1-x), which
        # yields a read whose close end is its last segment with the anchor to the right.
        first = ar < k[:, None]
        b = np.where(kind == 0, a + dsz, a)                    # deletion: skip dsz bases
        b = np.where(kind == 3, a - vsz, b)                    # tandem dup: jump back to the segment start
        b = np.where(kind == 2, a + vsz, b)                    # inversion: second part read backwards from a+vsz
        j2 = ar - k[:, None]                                   # offset inside the second segment
        ins = (kind == 1)
        j2 = np.where(ins[:, None], j2 - isz[:, None], j2)     # insertion: second segment starts after isz bases
        idx2 = np.where((kind == 2)[:, None], b[:, None] - 1 - j2, b[:, None] + j2)
        idxm = np.where(first, a[:, None] - k[:, None] + ar, idx2)
        inserted = ins[:, None] & (~first) & (j2 < 0)
        idxm = np.clip(idxm, 0, G - 1)
        plus = rng.random(m) < 0.5
        t = np.where(plus[:, None], ref[idxm], _COMP[ref[G - 1 - idxm]])
        inv2 = (kind == 2)[:, None] & (~first)
        t = np.where(inv2, _COMP[t], t)
        rnd_ins = _ACGT[rng.integers(0, 4, size=(m, L), dtype=np.uint8)]
        t = np.where(inserted, rnd_ins, t)
        # sequencing errors: substitutions to a different base, then sprinkled N
        err = rng.random((m, L)) < sub_rate
        shift = rng.integers(1, 4, size=(m, L), dtype=np.uint8)
        sub = _ACGT[(_CODE[t] + shift) & 3]
        t = np.where(err & (t != ord("N")), sub, t)
        t = np.where(rng.random((m, L)) < n_rate, np.uint8(ord("N")), t)
        slack = rng.integers(1, max(2, insert_size - L), size=m)
        g_first = a - k                                        # model coordinate of the read's first base
        pos_plus = g_first - slack                             # window [A, A+IS) must contain g_first
        pos_minus = (G - 1 - g_first) + slack                  # window [A-IS, A) must contain the mirrored base
        s_rc = _COMP[t[:, ::-1]]                               # UnmatchedSeq = RC(model read) on either strand
        seqs[c0:c0 + m] = s_rc
        md[c0:c0 + m] = np.where(plus, ord("+"), ord("-"))
        mpos[c0:c0 + m] = np.where(plus, pos_plus, pos_minus).astype(np.uint32)
        kinds[c0:c0 + m] = kind
    return ReadBatch(seqs, md, mpos, np.full(n, insert_size, dtype=np.int16), kinds)


def make_reads_torch(ref_t, n: int, read_len: int, seed: int, insert_size: int = 500, mix=(0.60, 0.25, 0.05, 0.10),
                     decoy_frac: float = 0.10, max_del: int = 1000, sub_rate: float = 0.01, n_rate: float = 0.001,
                     chunk: int = 1 << 20, lo=None, hi=None):
    """Same read model as make_reads(), generated with torch ops on ref_t's device (bench-sized batches:
    10 M x 150 bp in seconds on a GPU).  ref_t: uint8 ASCII tensor (unpadded).  Returns torch tensors
    (seq [n, L] uint8, matched_d [n] uint8, matched_rel_pos [n] int64, insert_size [n] int16, kind [n] uint8).
    The random stream differs from the numpy generator; the distribution is the same."""
    import torch
    dev = ref_t.device
    L, G = read_len, ref_t.numel()
    margin = 4 * insert_size + 3 * L + max_del + 4096
    lo = margin if lo is None else max(lo, margin)
    hi = G - margin if hi is None else min(hi, G - margin)
    if hi <= lo:
        raise ValueError("reference too small for the requested read model")
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    comp = torch.zeros(256, dtype=torch.uint8, device=dev)
    for a_, b_ in zip(b"ACGTN", b"TGCAN"):
        comp[a_] = b_
    code = torch.zeros(256, dtype=torch.int64, device=dev)
    code[ord("C")], code[ord("G")], code[ord("T")] = 1, 2, 3
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    cum = torch.cumsum(torch.tensor(mix, dtype=torch.float64, device=dev), 0)
    cum = cum / cum[-1]
    ar = torch.arange(L, device=dev, dtype=torch.int64)[None, :]
    seqs = torch.empty((n, L), dtype=torch.uint8, device=dev)
    md = torch.empty(n, dtype=torch.uint8, device=dev)
    mpos = torch.empty(n, dtype=torch.int64, device=dev)
    kinds = torch.empty(n, dtype=torch.uint8, device=dev)

    def rnd(*shape):
        return torch.rand(*shape, generator=gen, device=dev, dtype=torch.float64)

    def rint(lo_, hi_, *shape):
        return torch.randint(lo_, hi_, shape, generator=gen, device=dev, dtype=torch.int64)

    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        kind = torch.searchsorted(cum, rnd(m), right=True).clamp_(max=3)
        kind = torch.where(rnd(m) < decoy_frac, torch.full_like(kind, 4), kind)
        k = rint(20, L - 20 + 1, m)
        a = rint(lo, hi, m)
        dsz = torch.exp(rnd(m) * float(np.log(max_del))).to(torch.int64)
        isz = rint(1, 51, m)
        vsz = rint(50, max_del + 1, m)
        first = ar < k[:, None]
        b = torch.where(kind == 0, a + dsz, a)
        b = torch.where(kind == 3, a - vsz, b)
        b = torch.where(kind == 2, a + vsz, b)
        j2 = ar - k[:, None]
        ins = kind == 1
        j2 = torch.where(ins[:, None], j2 - isz[:, None], j2)
        idx2 = torch.where((kind == 2)[:, None], b[:, None] - 1 - j2, b[:, None] + j2)
        idxm = torch.where(first, a[:, None] - k[:, None] + ar, idx2).clamp_(0, G - 1)
        inserted = ins[:, None] & (~first) & (j2 < 0)
        plus = rnd(m) < 0.5
        t = torch.where(plus[:, None], ref_t[idxm], comp[ref_t[G - 1 - idxm].long()])
        inv2 = (kind == 2)[:, None] & (~first)
        t = torch.where(inv2, comp[t.long()], t)
        t = torch.where(inserted, acgt[rint(0, 4, m, L)], t)
        err = torch.rand(m, L, generator=gen, device=dev) < sub_rate
        sub = acgt[(code[t.long()] + rint(1, 4, m, L)) & 3]
        t = torch.where(err & (t != ord("N")), sub, t)
        t = torch.where(torch.rand(m, L, generator=gen, device=dev) < n_rate, torch.full_like(t, ord("N")), t)
        slack = rint(1, max(2, insert_size - L), m)
        g_first = a - k
        seqs[c0:c0 + m] = comp[t.flip(1).long()]
        md[c0:c0 + m] = torch.where(plus, torch.full_like(k, ord("+")), torch.full_like(k, ord("-"))).to(torch.uint8)
        mpos[c0:c0 + m] = torch.where(plus, g_first - slack, (G - 1 - g_first) + slack)
        kinds[c0:c0 + m] = kind.to(torch.uint8)
    return seqs, md, mpos, torch.full((n,), insert_size, dtype=torch.int16, device=dev), kinds


def make_discordant_pairs(n_events: int, seed: int, span: int = 60_000_000, background: int = 0, max_support: int = 40,
                          libraries=((100, 250), (150, 500)), both_mates: float = 0.8) -> np.ndarray:
    """Synthetic input of the read-pair clustering (row N2): what build_record_RP_Discovery (src/reader.cpp:982-1097)
    leaves in Reads_RP_Discovery for one chromosome -- PosA < PosB, strands, read length and insert size of the library.
    Events: two breakpoints with `support` pairs scattered within one insert size of them (most pairs twice: the BAM holds
    both mates and both records normalise to the same row); `background` isolated pairs.  -> structured array
    (oracle.rporacle.RP_DT layout), in BAM order (sorted by the position of the record that produced the row)."""
    rng = np.random.default_rng(seed)
    rows = []
    for _ in range(n_events):
        rl, isz = libraries[int(rng.integers(0, len(libraries)))]
        a = int(rng.integers(10_000, max(20_000, span - 2_000_000)))
        kind = rng.random()
        dist = int(rng.integers(50, 400)) if kind < 0.1 else int(rng.integers(800, 5000)) if kind < 0.7 else int(rng.integers(5000, 1_500_000))
        b = a + dist
        da, db = (b"+-", b"-+", b"++", b"--")[int(rng.integers(0, 4))]
        for _ in range(int(rng.integers(1, max_support + 1))):
            pa = a + int(rng.integers(-isz, isz))
            pb = b + int(rng.integers(-isz, isz))
            if pa >= pb:
                continue
            for _ in range(2 if rng.random() < both_mates else 1):
                rows.append((pa, pb, isz, rl, da, db))
    for _ in range(background):
        rl, isz = libraries[int(rng.integers(0, len(libraries)))]
        pa = int(rng.integers(10_000, max(20_000, span - 2_000_000)))
        pb = pa + int(rng.integers(1, 1_000_000))
        da, db = (b"+-", b"-+", b"++", b"--")[int(rng.integers(0, 4))]
        rows.append((pa, pb, isz, rl, da, db))
    out = np.zeros(len(rows), dtype=[("pos_a", "<u4"), ("pos_b", "<u4"), ("insert_size", "<u4"), ("read_length", "<i2"), ("da", "u1"), ("db", "u1")])
    for k, r in enumerate(rows):
        out[k] = r
    return out[rng.permutation(len(rows))]


def make_bam_like_bytes(n: int, seed: int) -> bytes:
    """n bytes that compress like BAM records do (for the BGZF inflater's tests and benchmark): per 200-byte record eight
    little-endian integers (refID, an increasing position, bin/mapq, flag, l_seq, mate fields), a read name with a running
    number, 50 bytes of 4-bit packed sequence and 100 quality bytes with runs.  Vectorised."""
    rng = np.random.default_rng(seed)
    k = n // 200 + 1
    rec = np.zeros((k, 200), dtype=np.uint8)
    core = np.zeros((k, 8), dtype="<i4")
    core[:, 1] = 10_000 + np.cumsum(rng.integers(0, 40, k))
    core[:, 2] = 0x12340000 | rng.integers(0, 61, k)
    core[:, 3] = rng.choice(np.array([99, 147, 83, 163, 73, 133], dtype=np.int32), k)
    core[:, 4] = 100
    core[:, 6] = core[:, 1] + rng.integers(200, 600, k)
    core[:, 7] = rng.integers(-600, 600, k)
    rec[:, :32] = core.view(np.uint8).reshape(k, 32)
    num = rng.integers(0, 10**9, k)
    name = np.frombuffer(b"read_000000000\0\0\0\0", dtype=np.uint8).copy()
    rec[:, 32:50] = name
    for d in range(9):
        rec[:, 45 - d] = 48 + (num // 10**d) % 10
    rec[:, 50:100] = rng.integers(0, 256, (k, 50), dtype=np.uint8)     # 4-bit bases: incompressible beyond the symbol skew
    q = np.repeat(rng.integers(2, 41, (k, 10), dtype=np.uint8), 10, axis=1)
    noise = rng.random((k, 100)) < 0.15
    q[noise] = rng.integers(2, 41, int(noise.sum()), dtype=np.uint8)
    rec[:, 100:200] = q + 33
    return rec.tobytes()[:n]


def make_interchr_pairs(n_events: int, seed: int, n_chr: int = 3, max_support: int = 30, background: int = 0):
    """Synthetic Reads_RP_Discovery_InterChr (row N2, inter-chromosomal variants): pairs whose mates lie on different
    chromosomes, stored from the read's side (ChrNameA = the read's chromosome), so one translocation shows up in both
    orientations -- (X, Y) rows from X's window and (Y, X) rows from Y's.  -> (rows RP_DT layout, chr_a, chr_b)"""
    rng = np.random.default_rng(seed)
    rows, ca, cb = [], [], []
    for _ in range(n_events):
        rl, isz = ((100, 250), (150, 500))[int(rng.integers(0, 2))]
        x, y = rng.choice(n_chr, 2, replace=False)
        px, py = int(rng.integers(10_000, 5_000_000)), int(rng.integers(10_000, 5_000_000))
        dx, dy = (b"+-", b"-+", b"++", b"--")[int(rng.integers(0, 4))]
        for _ in range(int(rng.integers(1, max_support + 1))):
            ax, ay = px + int(rng.integers(-isz, isz)), py + int(rng.integers(-isz, isz))
            for _ in range(2 if rng.random() < 0.5 else 1):          # exact copies (duplicates / re-reads)
                if rng.random() < 0.5:
                    rows.append((ax, ay, isz, rl, dx, dy)); ca.append(int(x)); cb.append(int(y))
                else:
                    rows.append((ay, ax, isz, rl, dy, dx)); ca.append(int(y)); cb.append(int(x))
    for _ in range(background):
        rl, isz = ((100, 250), (150, 500))[int(rng.integers(0, 2))]
        x, y = rng.choice(n_chr, 2, replace=False)
        d = (b"+-", b"-+", b"++", b"--")[int(rng.integers(0, 4))]
        rows.append((int(rng.integers(10_000, 5_000_000)), int(rng.integers(10_000, 5_000_000)), isz, rl, d[0], d[1])); ca.append(int(x)); cb.append(int(y))
    out = np.zeros(len(rows), dtype=[("pos_a", "<u4"), ("pos_b", "<u4"), ("insert_size", "<u4"), ("read_length", "<i2"), ("da", "u1"), ("db", "u1")])
    for k, r in enumerate(rows):
        out[k] = r
    perm = rng.permutation(len(rows))
    return out[perm], np.array(ca, dtype=np.int32)[perm], np.array(cb, dtype=np.int32)[perm]
