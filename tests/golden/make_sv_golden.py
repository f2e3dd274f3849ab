"""tests/golden/make_sv_golden.py -- regenerates the sv_*.npz fixtures FROM THE REFERENCE ITSELF.

For every search fixture of this directory (the synthetic batches and the COLO-829 slice, whose per-read UniquePoint lists
were dumped from the reference by make_golden.py) it runs the reference's seven SV drivers (src/pindel.cpp:1140-1164)
through oracle/_ref/libpindel_ref.so (oracle/ref_hook.cpp, pgref_classify) on the window parameters listed in CASES and
stores the dump: the box lists each driver hands to its reporter, the reads it saves for the next window and the
SPLIT_READ fields of every read after each driver.

    python tests/golden/make_sv_golden.py          # needs /root/reference and `make -C oracle ref`
"""
from __future__ import annotations

import gzip
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# name: (window_end fraction of the chromosome, region (start, end) fractions, driver mask, extra pindel args,
#        num_boxes override, every k-th read Used, every k-th read on another chromosome)
CASES = {
    "synth_L100_default": (0.6, (0.0, 1.0), 0x7F, [], None, 0, 0),
    "synth_L150_default": (1.0, (0.25, 0.66), 0x7F, [], None, 0, 0),
    "synth_L75_x4_u05": (0.8, (0.0, 1.0), 0x7F, ["-v", "10", "-M", "20"], 12000, 0, 0),
    "synth_L100_a2_m5_H10_e03": (1.0, (0.0, 1.0), 0b1000011, [], None, 7, 11),
    "synth_L250_default": (0.5, (0.0, 1.0), 0x7F, [], None, 0, 0),
    "colo_slice": (1.0, (0.0, 1.0), 0x7F, [], None, 0, 0),
}


def case_inputs(name):
    """-> (Case, classify inputs dict, window dict, fasta writer)"""
    from tests import fixtures, sv_helpers as S
    case = fixtures.load_colo_slice() if name == "colo_slice" else fixtures.load_synth(name[len("synth_"):])
    frac_end, (r0, r1), mask, extra, nb, k_used, k_other = CASES[name]
    seqs = S.final_seqs(case.blob, case.off, case.expected)
    ci = S.classify_inputs(case.expected, seqs, case.md, case.isz)
    n = len(ci["index"])
    if k_used:
        ci["rflags"][::k_used] |= 1
    if k_other:
        ci["rflags"][::k_other] |= 2
    comp = int(case.padded.size)  # Chromosome::getCompSize: the padded string
    bio = comp - 2 * 100000
    box = comp // 30000 or 1
    win = dict(window_end=int(bio * frac_end), region=(int(bio * r0), int(bio * r1)), driver_mask=mask,
               box_size=box, num_boxes=nb if nb else comp * 2 // box + 1)
    return case, ci, win, extra, seqs


def worker(name):
    from oracle import refhook
    case, ci, win, extra, _ = case_inputs(name)
    tmp = tempfile.mkdtemp()
    fa = os.path.join(tmp, "ref.fa")
    bio = case.padded[100000:case.padded.size - 100000 - 1]  # without the spacers and the repeated last base
    refhook.write_fasta(fa, [("c", bio)])
    args = []
    for k, v in case.settings.items():
        flag = {"max_range_index": "-x", "max_mismatch_rate": "-u", "additional_mismatch": "-a", "min_perfect_match": "-m",
                "min_close": "-H", "seq_error_rate": "-e", "sensitivity": "-E"}[k]
        args += [flag, str(v)]
    h = refhook.RefHook(fa, args + extra, workdir=tmp)
    assert h.chr_seq(0).size == case.padded.size and np.array_equal(h.chr_seq(0), case.padded)
    kw = {k: ci[k] for k in ("seq_blob", "seq_off", "matched_d", "insert_size", "rflags", "up_rows", "up_off", "n_close")}
    stream, _ = h.classify(**kw, window=(0, win["window_end"]), num_boxes=win["num_boxes"], box_size=win["box_size"],
                           driver_mask=win["driver_mask"], region=win["region"])
    np.savez_compressed(os.path.join(HERE, "sv_%s.npz" % name), stream=stream, n=len(ci["index"]),
                        window_end=win["window_end"], region=np.array(win["region"], np.int64), driver_mask=win["driver_mask"],
                        box_size=win["box_size"], num_boxes=win["num_boxes"], extra=np.array(extra, dtype="U8"),
                        min_inv=h.setting_int("MIN_IndelSize_Inversion"), min_matched=h.setting_int("Min_Num_Matched_Bases"),
                        seq_error_rate=h.setting_double("Seq_Error_Rate"), table=h.max_mismatch())
    from oracle import svoracle
    d = svoracle.parse_dump(stream, len(ci["index"]))
    print(name, {svoracle.DRIVERS[k]: (len(v["vars"]), len(v["future"])) for k, v in d.items()})


if __name__ == "__main__":
    if len(sys.argv) > 1:
        worker(sys.argv[1])
    else:
        for name in CASES:  # the reference keeps process-global state: one process per configuration
            subprocess.check_call([sys.executable, os.path.abspath(__file__), name])
