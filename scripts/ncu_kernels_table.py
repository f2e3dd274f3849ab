"""One row per captured kernel of an `ncu --set full` report (several kernels in one report) as a markdown table.

    python scripts/ncu_kernels_table.py gpurun_out/x.ncu-rep > profiles/x.md
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(txt)))
h, u = r[0], r[1]
cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp instr"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/instr"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
ix = [(h.index(k), n) for k, n in cols if k in h]
print("# ncu --set full: %s (digest; the report itself is not committed)\n" % rep)
print("| " + " | ".join("%s%s" % (n, (" [" + u[i] + "]") if u[i] else "") for i, n in ix) + " |")
print("|" + "---|" * len(ix))
for row in r[2:]:
    cells = []
    for i, n in ix:
        v = row[i]
        if n == "kernel":
            v = v.split("(")[0]
        else:
            try:
                f = float(v.replace(",", ""))
                v = ("%.0f" % f) if f >= 1000 or f == int(f) else ("%.2f" % f)
            except ValueError:
                pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
