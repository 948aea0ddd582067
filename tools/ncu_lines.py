#!/usr/bin/env python
"""Per-file / per-line aggregation of an `ncu --page source --csv --print-source cuda,sass` dump
(multi-file aware):  python tools/ncu_lines.py dump.csv [B] [top]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
hdr = rows[2]; samp = hdr.index('Warp Stall Sampling (All Samples)'); inst = hdr.index('Instructions Executed')
cur = None; lines = collections.defaultdict(lambda: [0, 0, ''])
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < len(hdr) or r[2] != '-' or not r[0].isdigit(): continue
    k = (cur, int(r[0])); lines[k][0] += float(r[samp] or 0); lines[k][1] += float(r[inst] or 0); lines[k][2] = r[1]
ts = sum(v[0] for v in lines.values()); ti = sum(v[1] for v in lines.values())
print(rows[1][1]); print('total samples', ts, 'warp-inst/instance', ti / B)
for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f[:16]:16s}{ln:5d} samp {100*v[0]/ts:5.1f}% inst {100*v[1]/ti:5.1f}% ({v[1]/B:6.0f})  {v[2].strip()[:105]}")
