"""Hot code footprint of a kernel from an `ncu --import-source on` report: how many bytes of SASS account for the executed
warp instructions (the L1.5 instruction cache of an SM holds 32 KB).

    python scripts/ncu_hot_footprint.py gpurun_out/x.ncu-rep
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None
ex = []
for r in rows:
    if r and r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        ex.append(int(r[hdr["Instructions Executed"]]))
    except ValueError:
        pass
tot = sum(ex)
print("SASS instructions: %d (%.1f KB), executed at least once: %d (%.1f KB)" % (len(ex), len(ex) * 16 / 1024, sum(1 for e in ex if e), sum(1 for e in ex if e) * 16 / 1024))
srt = sorted(ex, reverse=True)
acc = 0
marks = [0.5, 0.8, 0.9, 0.95, 0.99, 0.999]
mi = 0
for i, e in enumerate(srt):
    acc += e
    while mi < len(marks) and acc >= marks[mi] * tot:
        print("%5.1f %% of the executed instructions come from %5d SASS instructions = %5.1f KB" % (100 * marks[mi], i + 1, (i + 1) * 16 / 1024))
        mi += 1
