#!/usr/bin/env python
"""Opcode histogram of an address range of a cuobjdump -sass listing (static instruction mix of a loop body).
usage: sass_mix.py file.sass 0xLO 0xHI [more ranges ...]"""
import collections
import re
import sys

pat = re.compile(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)")
rows = []
for line in open(sys.argv[1]):
    m = pat.match(line)
    if m:
        rows.append((int(m.group(1), 16), m.group(2)))
args = sys.argv[2:]
FMAHEAVY = ("IMAD",)  # every IMAD.* form issues on the fmaheavy/fmalite pipes; WIDE forms take two slots
for lo, hi in zip(args[::2], args[1::2]):
    lo, hi = int(lo, 16), int(hi, 16)
    c = collections.Counter(op for a, op in rows if lo <= a < hi)
    tot = sum(c.values())
    wide = sum(v for k, v in c.items() if k.startswith("IMAD.WIDE"))
    imad_other = sum(v for k, v in c.items() if k.startswith("IMAD") and not k.startswith("IMAD.WIDE"))
    print(f"range {lo:#x}..{hi:#x}: {tot} instructions; IMAD.WIDE* {wide}; other IMAD* {imad_other}; pipe slots (2 per WIDE) {2 * wide + imad_other}")
    for k, v in c.most_common(18):
        print(f"   {v:6d} {k}")
