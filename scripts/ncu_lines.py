"""Per-source-line digest of an `ncu --set full --import-source on` report: where the warp instructions of a kernel go.

    python scripts/ncu_lines.py gpurun_out/x.ncu-rep [reads_in_launch] [top_n]

Prints, per (file, line) of the CUDA sources, the executed warp instructions (absolute, share, per read) and the stall
samples, largest first, plus the totals per source function range given in RANGES below (pg_warp.cuh's phases).
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
reads = int(sys.argv[2]) if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40

txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
per = collections.OrderedDict()
cur_file = None
hdr = None
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        # two "Source" columns: the first is the CUDA line, the second the SASS text
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] == "":          # a SASS row under the current CUDA line
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    n = num(r[hdr["Instructions Executed"]])
    s = num(r[hdr["# Samples"]])
    if n or s:
        k = (cur_file, line)
        a = per.setdefault(k, [0, 0, r[1].strip()])
        a[0] += n
        a[1] += s
tot = sum(v[0] for v in per.values()) or 1
tots = sum(v[1] for v in per.values()) or 1
print("total warp instructions %d%s, stall samples %d" % (tot, (" (%.0f per read)" % (tot / reads)) if reads else "", tots))
print("\n%-22s %14s %7s %9s %7s  source" % ("file:line", "warp instr", "share", "per read", "samples"))
for (f, l), (n, s, src) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s %14d %6.2f%% %9s %6.2f%%  %s" % ("%s:%d" % (f, l), n, 100.0 * n / tot, ("%.1f" % (n / reads)) if reads else "-", 100.0 * s / tots, src[:90]))
byfile = collections.Counter()
for (f, l), (n, s, src) in per.items():
    byfile[f] += n
print("\nby file:")
for f, n in byfile.most_common():
    print("  %-28s %6.2f%%" % (f, 100.0 * n / tot))
# pg_warp.cuh line ranges -> phases (kept in step with the file by hand; the digest names the ranges it used)
if len(sys.argv) > 4:
    rng = []
    for spec in sys.argv[4:]:
        name, a, b = spec.split(":")
        rng.append((name, int(a), int(b)))
    print("\nby pg_warp.cuh line range:")
    for name, a, b in rng:
        n = sum(v[0] for (f, l), v in per.items() if f == "pg_warp.cuh" and a <= l <= b)
        print("  %-28s %5d-%-5d %6.2f%%%s" % (name, a, b, 100.0 * n / tot, ("  %.0f per read" % (n / reads)) if reads else ""))
