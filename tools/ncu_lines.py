#!/usr/bin/env python3
"""Join an ncu SASS source page (csv) with nvdisasm -gi line info: samples / instructions per
top-level source line (the outermost 'inlined at' frame in zw_search.cuh etc.).
usage: ncu_lines.py <ncu-rep> <kernel substring> <cubin-disasm-gi.txt> <mangled kernel name>"""
import csv, re, subprocess, sys, collections

rep, ksub, disasm, mangled = sys.argv[1:5]
# --- disasm: offset -> (innermost file:line, outermost file:line)
off2 = {}
active = False
frames = []      # frames of the pending instruction, innermost first
fresh = True
for ln in open(disasm):
    if ln.startswith(".text."):
        active = ln.strip().rstrip(":") == ".text." + mangled
        continue
    if not active:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            frames = []
            fresh = False
        frames.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
    if m:
        fresh = True
        if frames:
            # innermost, and the frame directly below the kernel body (second outermost)
            off2[int(m.group(1), 16)] = (frames[0], frames[-2] if len(frames) >= 2 else frames[-1])
# --- ncu sass page
if rep.endswith(".gz"):
    import gzip
    out = gzip.open(rep, "rt").read()
elif rep.endswith(".csv"):
    out = open(rep).read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern = None; hdr = None; base = None
agg_out = collections.Counter(); agg_in = collections.Counter(); inst_out = collections.Counter(); inst_in = collections.Counter()
tot_s = tot_i = 0
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        kern = r[1]; hdr = None; base = None
        continue
    if kern is None or ksub not in kern:
        continue
    if hdr is None:
        hdr = r; ia = hdr.index("Address"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
        continue
    addr = int(r[ia], 16)
    if base is None:
        base = addr
    off = addr - base
    s = int(r[isamp] or 0); e = int(r[iex] or 0)
    inner, outer = off2.get(off, (("?", 0), ("?", 0)))
    agg_out[outer] += s; agg_in[inner] += s; inst_out[outer] += e; inst_in[inner] += e
    tot_s += s; tot_i += e
print("kernel", ksub, "samples", tot_s, "warp-inst", tot_i)
print("--- by outermost line (samples%, inst%)")
for k, v in agg_out.most_common(45):
    print("%-22s %6d  %5.1f%%   inst %5.1f%%" % ("%s:%d" % k, v, 100.0 * v / tot_s, 100.0 * inst_out[k] / tot_i))
print("--- by innermost line")
for k, v in agg_in.most_common(25):
    print("%-22s %6d  %5.1f%%   inst %5.1f%%" % ("%s:%d" % k, v, 100.0 * v / tot_s, 100.0 * inst_in[k] / tot_i))
