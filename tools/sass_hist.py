#!/usr/bin/env python3
"""Opcode histogram per kernel from `cuobjdump -sass` output (stdin or file)."""
import re, sys, collections
fn = None
hist = collections.defaultdict(collections.Counter)
for line in (open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin):
    m = re.search(r'Function : (\S+)', line)
    if m:
        fn = m.group(1); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and fn:
        op = m.group(1)
        base = op.split('.')[0]
        key = base
        if base == 'IMAD':
            key = 'IMAD.WIDE' if '.WIDE' in op else ('IMAD.HI' if '.HI' in op else ('IMAD.MOV' if '.MOV' in op else ('IMAD.IADD' if '.IADD' in op else ('IMAD.SHL' if '.SHL' in op else 'IMAD'))))
        hist[fn][key] += 1
pat = sys.argv[2] if len(sys.argv) > 2 else ''
for f, h in hist.items():
    if pat and pat not in f: continue
    tot = sum(h.values())
    print(f"== {f}  total {tot}")
    print("   " + "  ".join(f"{k}:{v}" for k, v in h.most_common()))
