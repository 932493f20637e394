#!/usr/bin/env python3
"""registers / spills per kernel from the *.ptxas.log files the Makefile writes (nvcc -Xptxas -v)"""
import re, subprocess, sys, os
d = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "keyhunt_b200", "csrc")
for f in sorted(os.listdir(d)):
    if not f.endswith(".ptxas.log"):
        continue
    t = open(os.path.join(d, f)).read()
    for m in re.finditer(r"Compiling entry function '(\w+)' for 'sm_100a'\n(?:.*\n)*?ptxas info\s+: Function properties for \1\n\s+(.*)\nptxas info\s+: Used (\d+) registers", t):
        dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(.*", "", dem)
        print("%-12s %-48s regs %3s | %s" % (f.split(".")[0], dem[:48], m.group(3), m.group(2)))
