"""Loaders for the committed golden vectors (tests/golden/, produced by make_golden.py from the reference itself)."""
from __future__ import annotations

import gzip
import hashlib
import os

import numpy as np

from oracle import refhook
from pindel_b200 import synth
from tests.golden.make_golden import SYNTH

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SETTING_FLAGS = {"-x": ("max_range_index", int), "-u": ("max_mismatch_rate", float), "-a": ("additional_mismatch", int),
                 "-m": ("min_perfect_match", int), "-H": ("min_close", int), "-e": ("seq_error_rate", float),
                 "-E": ("sensitivity", float)}


def settings_from_args(args):
    out = {}
    for flag, val in zip(args[0::2], args[1::2]):
        name, conv = SETTING_FLAGS[flag]
        out[name] = conv(val)
    return out


class Case:
    """One golden case: inputs (padded chromosome, SoA batch arrays), settings, and the reference's per-read results."""

    def __init__(self, name, padded, blob, off, md, pos, isz, settings, expected):
        self.name, self.padded, self.blob, self.off, self.md, self.pos, self.isz = name, padded, blob, off, md, pos, isz
        self.settings, self.expected = settings, expected
        self.n = len(off) - 1

    def kwargs(self, lo=0, hi=None):
        hi = self.n if hi is None else min(hi, self.n)
        b0 = int(self.off[lo])
        return dict(seq_blob=self.blob[b0:int(self.off[hi])], seq_off=(self.off[lo:hi + 1] - b0).astype(np.uint32),
                    matched_d=self.md[lo:hi], matched_rel_pos=self.pos[lo:hi], insert_size=self.isz[lo:hi])


def load_synth(name) -> Case:
    ref_bases, ref_seed, n_runs, n, L, seed, extra, gkw = SYNTH[name]
    z = np.load(os.path.join(GOLD, "synth_%s.npz" % name))
    ref = synth.make_reference(ref_bases, ref_seed, n_runs=n_runs)
    batch = synth.make_reads(ref, n, L, seed=seed, **gkw)
    digest = np.frombuffer(hashlib.sha256(batch.seq.tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, z["seq_digest"]), "the seeded generator no longer reproduces the fixture inputs"
    padded = synth.pad_reference(ref)
    # Genome::loadChromosome appends the last base once more at end of file (src/pindel.cpp:287-300): mirror it
    padded = reference_memory_image(ref)
    assert padded.size == int(z["chr_len"])
    blob, off = batch.blob()
    return Case(name, padded, blob, off, batch.matched_d, batch.matched_rel_pos, batch.insert_size,
                settings_from_args([str(a) for a in z["args"]]), refhook.parse_stream(z["stream"], n))


def reference_memory_image(ref: np.ndarray) -> np.ndarray:
    """The chromosome string as the reference holds it after loading a FASTA that ends in a newline: 100 000 N, the
    sequence, its LAST BASE REPEATED ONCE (the do/while in Genome::loadChromosome re-appends the last character when
    `>>` hits end of file, src/pindel.cpp:287-300), 100 000 N."""
    out = np.full(ref.size + 1 + 2 * synth.SPACER, ord("N"), dtype=np.uint8)
    out[synth.SPACER:synth.SPACER + ref.size] = ref
    out[synth.SPACER + ref.size] = ref[-1]
    return out


def load_edge() -> Case:
    z = np.load(os.path.join(GOLD, "edge_cases.npz"))
    ref = synth.make_reference(300_000, 106, n_runs=[(150_000, 700)])
    padded = reference_memory_image(ref)
    assert padded.size == int(z["chr_len"])
    n = len(z["off"]) - 1
    return Case("edge_cases", padded, z["blob"], z["off"], z["md"], z["pos"], z["isz"], {}, refhook.parse_stream(z["stream"], n))


def load_colo_slice() -> Case:
    z = np.load(os.path.join(GOLD, "colo_slice_stream.npz"))
    fa = gzip.open(os.path.join(GOLD, "colo_slice.fa.gz")).read()
    seq = np.frombuffer(fa[fa.index(b"\n") + 1:].replace(b"\n", b""), dtype=np.uint8)
    padded = reference_memory_image(seq)
    assert padded.size == int(z["chr_len"])
    lines = gzip.open(os.path.join(GOLD, "colo_slice_reads.txt.gz")).read().decode().split("\n")
    seqs, md, pos, isz = [], [], [], []
    for i in range(0, len(lines) - 2, 3):
        seqs.append(lines[i + 1].encode())
        f = lines[i + 2].split()
        md.append(ord(f[0])); pos.append(int(f[2])); isz.append(int(f[4]))
    off = np.zeros(len(seqs) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(s) for s in seqs])
    n = int(z["n"])
    assert n == len(seqs)
    return Case("colo_slice", padded, np.frombuffer(b"".join(seqs), dtype=np.uint8), off, np.array(md, np.uint8),
                np.array(pos, np.uint32), np.array(isz, np.int16), {}, refhook.parse_stream(z["stream"], n))


ALL_SYNTH = list(SYNTH)
