#!/usr/bin/env python
"""Top warp-stall sampling sites (SASS) per kernel from an `ncu -i X.ncu-rep --page source --csv | gzip` dump.
usage: python scripts/ncu_top_stalls.py profiles/NAME_source.csv.gz [N] > profiles/NAME_top_stall_sites.txt"""
import collections
import csv
import gzip
import io
import sys

rows = list(csv.reader(io.TextIOWrapper(gzip.open(sys.argv[1]), newline="")))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 14
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
print(f"# top warp-stall sampling sites (SASS) of the kernels in {sys.argv[1]} (source page of the ncu --set full capture);")
print("# share = samples at the instruction / all samples of the kernel; the reason columns are the sampler's own (samples per reason)")
seen = set()
for n, s in enumerate(starts):
    end = starts[n + 1] if n + 1 < len(starts) else len(rows)
    name, hdr, body = rows[s][1], rows[s + 1], [r for r in rows[s + 2:end] if len(r) > 5 and r[2].isdigit()]
    tot = sum(int(r[2]) for r in body)
    if (name, tot) in seen:  # the dump repeats every kernel (one table per view)
        continue
    seen.add((name, tot))
    reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    print(f"\n== {name[:110]}\n   {len(body)} SASS instructions, {tot} samples")
    for r in sorted(body, key=lambda r: -int(r[2]))[:top_n]:
        why = sorted(((int(r[i]), h[6:]) for i, h in reasons if i < len(r) and r[i].isdigit() and int(r[i])), reverse=True)[:2]
        print(f"   {int(r[2]) / tot * 100:5.2f} %  {r[1].strip()[:64]:64s} {' '.join(f'{h}={v}' for v, h in why)}")
    by_op = collections.Counter()
    for r in body:
        op = r[1].split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "")
        by_op[op] += int(r[2])
    print("   by opcode: " + ", ".join(f"{k} {v / tot * 100:.1f}%" for k, v in by_op.most_common(8)))
    if reasons:
        by_reason = collections.Counter()
        for r in body:
            for i, h in reasons:
                if i < len(r) and r[i].isdigit():
                    by_reason[h[6:]] += int(r[i])
        rt = sum(by_reason.values()) or 1
        print("   by reason: " + ", ".join(f"{k} {v / rt * 100:.1f}%" for k, v in by_reason.most_common(8)))
