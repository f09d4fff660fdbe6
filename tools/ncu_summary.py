#!/usr/bin/env python3
"""Summarise an ncu --set full report into profiles/<name>_summary.md and profiles/<name>_traffic.json.
usage: python tools/ncu_summary.py <ncu-rep | raw-page csv> r1 <pixels per launch>"""
import csv, json, subprocess, sys, collections

rep, name, pixels = sys.argv[1], sys.argv[2], float(sys.argv[3])
if rep.endswith(".csv"):  # the raw page exported on the GPU box (ncu -i X.ncu-rep --page raw --csv)
    out = open(rep).read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
M = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs/thread"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
     ("smsp__inst_executed.sum", "warp instructions"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
     ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe % of peak"),
     ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA/IMAD pipe % of peak"),
     ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe % of peak"),
     ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instr (divergence)"),
     ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle"),
     ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
     ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
     ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %")]
seen = collections.OrderedDict()
for r in rows[2:]:
    k = r[col["Kernel Name"]]
    if k not in seen:
        seen[k] = r
md = ["# ncu --set full summary (%s)" % name, "",
      "Source: `%s` (`ncu --set full --clock-control none --import-source on`), one launch per kernel, "
      "workload = %d pixels per launch (batch of 768x512 images, q75 m4).  Times under ncu are cold-cache and "
      "serialised: compare shares, not absolutes." % (rep.split("/")[-1], pixels), ""]
traffic = {}
def num(r, m):
    try:
        return float(r[col[m]].replace(",", ""))
    except Exception:
        return None
for k, r in seen.items():
    md += ["## `%s`" % k, "", "| metric | value |", "|---|---|"]
    for m, label in M:
        if m in col:
            md.append("| %s | %s %s |" % (label, r[col[m]], units[col[m]]))
    tot = sum(float(r[i] or 0) for h, i in col.items() if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued"))
    st = sorted(((float(r[i] or 0), h[32:]) for h, i in col.items() if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")), reverse=True)
    md.append("| warp-state samples (top) | " + ", ".join("%s %.1f%%" % (h.lstrip("_"), 100 * v / tot) for v, h in st[:6] if tot) + " |")
    rd, wr, du = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum"), num(r, "gpu__time_duration.sum")
    if rd is not None and wr is not None:
        ur, uw = units[col["dram__bytes_read.sum"]], units[col["dram__bytes_write.sum"]]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        b = rd * scale.get(ur, 1) + wr * scale.get(uw, 1)
        traffic[k] = {"dram_bytes_per_launch": b, "dram_bytes_per_pixel": b / pixels}
        md.append("| DRAM bytes / pixel | %.2f |" % (b / pixels))
    md.append("")
open("profiles/%s_summary.md" % name, "w").write("\n".join(md) + "\n")
json.dump({"pixels_per_launch": pixels, "kernels": traffic}, open("profiles/%s_traffic.json" % name, "w"), indent=1)
print("wrote profiles/%s_summary.md" % name)
