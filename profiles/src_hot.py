#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per source line:
   python profiles/src_hot.py dump.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[2]
ci = {h: i for i, h in enumerate(hdr)}
samp = hdr.index('Warp Stall Sampling (All Samples)')
inst = hdr.index('Instructions Executed')
conf = hdr.index('L1 Wavefronts Shared Excessive') if 'L1 Wavefronts Shared Excessive' in hdr else None
lines = {}
for r in rows[3:]:
    if len(r) < len(hdr): continue
    if r[0] not in ('', 'Line No') and r[2] == '-':
        ln = int(r[0]); src = r[1]
        d = lines.setdefault(ln, dict(src=src, samp=0, inst=0, conf=0))
        def num(x):
            try: return float(x)
            except: return 0.0
        d['samp'] += num(r[samp]); d['inst'] += num(r[inst])
        if conf is not None: d['conf'] += num(r[conf])
tot_s = sum(d['samp'] for d in lines.values()); tot_i = sum(d['inst'] for d in lines.values())
print(f"total samples {tot_s:.0f} total warp-insts {tot_i:.0f}")
for ln, d in sorted(lines.items(), key=lambda kv: -kv[1]['samp'])[:top]:
    print(f"{ln:5d} samp {100*d['samp']/tot_s:5.1f}% inst {100*d['inst']/tot_i:5.1f}% excess_wf {d['conf']:.3g}  {d['src'].strip()[:110]}")
