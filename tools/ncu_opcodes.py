#!/usr/bin/env python3
"""Opcode-class breakdown per kernel from an ncu report's SASS source page (Instructions Executed per SASS
line, summed by opcode): which instruction classes a kernel's issue slots go to.
usage: python tools/ncu_opcodes.py <ncu-rep> <out.md> [kernel substring ...]"""
import collections
import csv
import re
import subprocess
import sys

rep, out_md = sys.argv[1], sys.argv[2]
want = sys.argv[3:]
if rep.endswith(".gz"):
    import gzip
    txt = gzip.open(rep, "rt").read()
elif rep.endswith(".csv"):
    txt = open(rep).read()
else:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))

CLASSES = [
    ("shuffle (SHFL)", r"^SHFL"), ("vote/redux/match (VOTE, REDUX, MATCH)", r"^(VOTE|REDUX|MATCH|WARPSYNC|NANOSLEEP)"),
    ("shared load (LDS)", r"^LDS"), ("shared store / atomic (STS, ATOMS)", r"^(STS|ATOMS)"),
    ("local memory (LDL, STL)", r"^(LDL|STL)"), ("global/const load (LDG, LD, LDC, LDCU, ULDC)", r"^(LDG|LD\b|LD\.|LDC|LDCU|ULDC)"),
    ("global store / atomic (STG, ST, ATOMG, RED, MEMBAR, CCTL)", r"^(STG|ST\b|ST\.|ATOMG|ATOM|RED|MEMBAR|CCTL|ERRBAR)"),
    ("integer multiply-add (IMAD, IMUL, IDP)", r"^(IMAD|IMUL|IDP|UIMAD)"),
    ("integer add / logic / shift / min-max (IADD3, LOP3, SHF, LEA, IMNMX, IABS, VIADD, VIMNMX, VABSDIFF, PRMT, SGXT, BMSK, FLO, POPC, BREV)",
     r"^(IADD|VIADD|LOP|SHF|SHL|SHR|LEA|IMNMX|VIMNMX|IABS|VABSDIFF|PRMT|SGXT|BMSK|FLO|POPC|BREV|UIADD|ULOP|USHF|ULEA|UFLO|UPOPC|UPRMT|USGXT|UBMSK|UBREV|VIMNMX3)"),
    ("compare / select / predicate (ISETP, SEL, PLOP3, P2R, R2P)", r"^(ISETP|SEL|PLOP|P2R|R2P|UISETP|USEL|UPLOP|FSETP|ICMP|UP2UR)"),
    ("move / convert (MOV, I2F, F2I, CS2R, S2R, R2UR, ...)", r"^(MOV|UMOV|I2F|F2I|I2I|F2F|CS2R|S2R|S2UR|R2UR|UR2R|I2FP|F2FP|MOVM)"),
    ("branch / barrier / control (BRA, BSSY, BSYNC, BAR, CALL, RET, EXIT, ...)", r"^(BRA|BRX|JMP|BSSY|BSYNC|BAR|CALL|RET|EXIT|YIELD|NOP|BREAK|BMOV|WARPSYNC|DEPBAR|ACQBULK|ENDCOLLECTIVE|UBRA)"),
    ("floating point (FADD, FMUL, FFMA, MUFU, ...)", r"^(FADD|FMUL|FFMA|MUFU|FMNMX|HADD|HMUL|HFMA|DADD|DMUL|DFMA|FSEL|FCHK)"),
]


def opcode_of(src):
    s = src.strip()
    s = re.sub(r"^@!?U?P\w+\s+", "", s)  # predicate prefix
    return s.split()[0].rstrip(";") if s else "?"


kern = None
hdr = None
data = collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        kern = r[1]
        hdr = None
        continue
    if kern is None or not r:
        continue
    if r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or "Instructions Executed" not in hdr:
        continue
    d = data.setdefault(kern, {"ops": collections.Counter(), "thr": collections.Counter(), "seen": set()})
    if r[0] in d["seen"]:
        continue  # the page lists every kernel twice (SASS and source views)
    d["seen"].add(r[0])
    try:
        n = int(r[hdr["Instructions Executed"]])
        t = int(r[hdr["Thread Instructions Executed"]])
    except ValueError:
        continue
    op = opcode_of(r[hdr["Source"]])
    d["ops"][op] += n
    d["thr"][op] += t

md = ["# Opcode-class breakdown (warp-level instructions executed)", "",
      "Source: `%s`, `ncu --page source` (Instructions Executed per SASS line, summed by opcode).  One launch per kernel." % rep.split("/")[-1], ""]
for k, d in data.items():
    if want and not any(w in k for w in want):
        continue
    tot = sum(d["ops"].values())
    thr = sum(d["thr"].values())
    if tot == 0:
        continue
    md += ["## `%s`" % k, "", "%.3f G warp instructions, %.2f active threads per instruction." % (tot / 1e9, thr / tot), "",
           "| class | warp instr | share |", "|---|---|---|"]
    left = collections.Counter(d["ops"])
    for name, pat in CLASSES:
        n = 0
        for op in list(left):
            if re.match(pat, op):
                n += left.pop(op)
        if n:
            md.append("| %s | %.3f G | %.1f %% |" % (name, n / 1e9, 100.0 * n / tot))
    rest = sum(left.values())
    if rest:
        md.append("| other (%s) | %.3f G | %.1f %% |" % (", ".join(op for op, _ in left.most_common(6)), rest / 1e9, 100.0 * rest / tot))
    md += ["", "Top opcodes: " + ", ".join("%s %.1f %%" % (op.split(".")[0] if False else op, 100.0 * n / tot) for op, n in d["ops"].most_common(14)), ""]
open(out_md, "w").write("\n".join(md) + "\n")
print("wrote", out_md)
