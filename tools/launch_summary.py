#!/usr/bin/env python
"""Summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: per-kernel totals, and the launches
of one headline step (presolve kernel followed by the kernels that find their lists empty) with their shares.
   python tools/launch_summary.py profiles/r02b_ncu_launches_bench.csv"""
import collections, csv, re, sys
import numpy as np
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
short = lambda n: re.sub(r"\(.*", "", n.replace("void ", "").replace("cmpc::", ""))[:60]
L = [(short(r[4]), float(r[14]) / 1e3) for r in rows]
agg = collections.OrderedDict()
for n, us in L:
    a = agg.setdefault(n, [0, 0.0, 0.0]); a[0] += 1; a[1] += us; a[2] = max(a[2], us)
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'max us':>10s}")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:60s} {a[0]:8d} {a[1]:12.1f} {a[2]:10.1f} {100 * a[1] / tot:6.1f}%")
# headline steps: a presolve launch > 60 us followed by riccati / solve / ripm launches all < 12 us
steps = []
i = 0
while i < len(L):
    if L[i][0].startswith("cmpc_presolve_kernel") and L[i][1] > 60:
        j = i + 1; grp = [L[i]]
        while j < len(L) and L[j][0].startswith(("cmpc_riccati", "cmpc_solve", "cmpc_ripm")) and L[j][1] < 12:
            grp.append(L[j]); j += 1
        if len(grp) == 4 and j - i == 4: steps.append(grp)
        i = j
    else:
        i += 1
if steps:
    # the same launch sequence also appears in the end-to-end leg, where the kernel reads its inputs over the host link:
    # keep the device-resident steps (presolve launch within 1.3x of the fastest one)
    fastest = min(s_[0][1] for s_ in steps)
    steps = [s_ for s_ in steps if s_[0][1] <= 1.3 * fastest]
    print(f"\nheadline step (config 2, 4096 trot instances, every instance settled by the presolve): the launches of one step, in order (median of {len(steps)} such steps)")
    med = [float(np.median([s[k][1] for s in steps])) for k in range(4)]
    for k in range(4):
        print(f"  {steps[0][k][0]:58s} {med[k]:8.2f} us  {100 * med[k] / sum(med):5.1f}% of the step")
    print(f"  step total {sum(med):.1f} us under ncu (cold-cache, serialised launches: the SHARE is what compares with bench.py's device-timed step)")
