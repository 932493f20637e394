#!/usr/bin/env python3
"""Executed (dynamic) opcode mix of the first kernel in an .ncu-rep captured with --set full --import-source on:
   python tools/ncu_opmix.py report.ncu-rep [out.txt]
Sums the per-instruction 'Instructions Executed' column of the SASS source page by opcode and by pipe class."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Instructions Executed" in c for c in r):
        hdr = r; rows = rows[i + 1:]; break
if hdr is None:
    sys.exit("no source page in %s (needs --set full --import-source on)" % rep)
si = hdr.index("Source")
ei = [i for i, c in enumerate(hdr) if c.strip() == "# Instructions Executed" or c.strip() == "Instructions Executed"][0]
ti = [i for i, c in enumerate(hdr) if "Thread Instructions Executed" in c]
ALU = {"IADD3", "LOP3", "SHF", "PRMT", "LEA", "SEL", "ISETP", "PLOP3", "VIADD", "IABS", "IMNMX", "VIMNMX", "FLO", "POPC", "BREV", "P2R", "R2P", "SGXT", "BMSK", "FSETP", "FSEL", "FMNMX"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2"}
ops, pipes, tot = collections.Counter(), collections.Counter(), 0
for r in rows:
    if len(r) <= max(si, ei):
        continue
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", r[si])
    if not m:
        continue
    try:
        n = int(float(r[ei]))
    except ValueError:
        continue
    base, suf = m.group(1), m.group(2)
    key = base
    if base == "IMAD":
        key = "IMAD.WIDE" if ".WIDE" in suf else ("IMAD.HI" if ".HI" in suf else ("IMAD.MOV" if ".MOV" in suf else ("IMAD.IADD" if ".IADD" in suf else ("IMAD.SHL" if ".SHL" in suf else "IMAD"))))
    ops[key] += n; tot += n
    pipes["alu" if base in ALU else ("fma" if base in FMA else ("lsu" if base in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "LD", "ST", "ATOMG", "REDG", "RED", "ATOM", "LDC") else "other"))] += n
out = ["executed warp instructions by opcode (source page of %s), total %d" % (rep, tot)]
for k, v in ops.most_common(40):
    out.append("   %-12s %14d  %5.1f %%" % (k, v, 100.0 * v / max(1, tot)))
out.append("by pipe class: " + "  ".join("%s %.1f %%" % (k, 100.0 * v / max(1, tot)) for k, v in pipes.most_common()))
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
