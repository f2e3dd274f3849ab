"""Turn an `ncu --set full --import-source on` report into the short text summary committed under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_search_kernel.md [reads_in_launch]
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
reads = int(sys.argv[3]) if len(sys.argv) > 3 else None


def page(name):
    txt = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(txt)))


raw = page("raw")
hdr, units, vals = raw[0], raw[1], raw[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_allocated", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_tma_ld.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__warps_eligible.avg.per_cycle_active"]
stalls = sorted(((float(v[0]), k) for k, v in d.items() if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", k)), reverse=True)

src = page("source")
kname = src[0][1] if src and len(src[0]) > 1 else "?"
sh = src[1]
ix = {h: i for i, h in enumerate(sh)}
byop, samp, tot, tsamp = collections.Counter(), collections.Counter(), 0, 0
lines = []
for r in src[2:]:
    if len(r) < len(sh):
        continue
    s = r[ix["Source"]].strip()
    n = int(r[ix["Instructions Executed"]] or 0)
    sm = int(r[ix["# Samples"]] or 0)
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", s)
    op = m.group(2).split(".")[0] if m else s
    byop[op] += n; samp[op] += sm; tot += n; tsamp += sm
    lines.append((sm, n, s))

with open(out, "w") as f:
    f.write("# ncu --set full summary: %s\n\nreport: %s (not committed; this is its digest)\n\n" % (kname, rep))
    f.write("| metric | value | unit |\n|---|---|---|\n")
    for k in keys:
        if k in d:
            f.write("| %s | %s | %s |\n" % (k, d[k][0], d[k][1]))
    if reads:
        t_ms = float(d["gpu__time_duration.sum"][0].replace(",", ""))
        if d["gpu__time_duration.sum"][1] == "us":
            t_ms /= 1e3
        f.write("\nreads in this launch: %d -> %.2f M reads/s under the profiler (not a bench number), %.0f warp-instructions/read, "
                "%.1f DRAM bytes/read\n" % (reads, reads / t_ms / 1e3, float(d["smsp__inst_executed.sum"][0].replace(",", "")) / reads,
                                           (float(d["dram__bytes_read.sum"][0].replace(",", "")) * (1e6 if d["dram__bytes_read.sum"][1] == "Mbyte" else 1e9 if d["dram__bytes_read.sum"][1] == "Gbyte" else 1) +
                                            float(d["dram__bytes_write.sum"][0].replace(",", "")) * (1e6 if d["dram__bytes_write.sum"][1] == "Mbyte" else 1e9 if d["dram__bytes_write.sum"][1] == "Gbyte" else 1)) / reads))
    f.write("\n## warp stall reasons (warps per issue-active cycle)\n\n")
    for v, k in stalls[:8]:
        f.write("- %s: %.3f\n" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    f.write("\n## executed SASS by opcode (share of warp instructions / share of stall samples)\n\n")
    for op, n in byop.most_common(22):
        f.write("- %-8s %5.2f%% / %5.2f%%\n" % (op, 100.0 * n / max(tot, 1), 100.0 * samp[op] / max(tsamp, 1)))
    f.write("\n## 25 SASS lines with the most stall samples\n\n```\n")
    for sm, n, s in sorted(lines, reverse=True)[:25]:
        f.write("%7d samples  %12d exec  %s\n" % (sm, n, s))
    f.write("```\n")
print("wrote", out)
