#!/usr/bin/env python
"""Per-phase share of stall samples and warp instructions of the solve kernel from an
`ncu --page source --csv --print-source cuda,sass` dump:  python profiles/phase_breakdown.py dump.csv [B]"""
import csv, sys, os
rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
hdr = rows[2]; samp = hdr.index('Warp Stall Sampling (All Samples)'); inst = hdr.index('Instructions Executed')
src = open(os.path.join(os.path.dirname(__file__), '..', 'cheeta-mpc_b200', 'csrc', 'cmpc_device.cuh')).read().split('\n')
def find(s):
    for i, l in enumerate(src):
        if s in l: return i + 1
    return 10**6
marks = [('group prims/layout', 1), ('tile helpers+potrf4', find('4x4 tile kernels')), ('chol', find('bool chol_bc4(')),
         ('fwd solve', find('void chol_fwd_bc4(')), ('bwd solve', find('void chol_bwd_bc4(')),
         ('symv', find('void symv_bc4(')), ('copy H', find('void copy_mat')), ('pyramid+polish helpers', find('friction pyramid rows')),
         ('prologue/inputs', find('cmpc_solve_kernel(const DevConfig')), ('build H,g', find('// lever arms r = des_foot_pos')),
         ('start point', find('// ---- strictly feasible start')), ('residual+conv', find('// ---- residuals (M holds')),
         ('polish', find('if (cfg.polish && ready && npolish < 3)')), ('M=H+D, pred rhs', find("// ---- M = H + C' diag")),
         ('pred/corr vector ops', find('double tmax = 0.0, sigma = 0.0;')), ('outputs', find('// ---- outputs')), ('end', 10**7)]
agg = [[0, 0] for _ in marks]; ts = ti = 0
for r in rows[3:]:
    if len(r) < len(hdr) or r[2] != '-' or not r[0].isdigit(): continue
    ln = int(r[0]); s = float(r[samp] or 0); i = float(r[inst] or 0)
    k = max(j for j, (n, l) in enumerate(marks) if l <= ln)
    agg[k][0] += s; agg[k][1] += i; ts += s; ti += i
print(f"total stall samples {ts:.0f}, warp instructions {ti:.0f} ({ti/B:.0f} per instance)")
for (n, l), (s, i) in zip(marks, agg):
    if n != 'end': print(f"{n:24s} samples {100*s/ts:5.1f}%  insts {100*i/ti:5.1f}% ({i/B:7.0f}/instance)")
