#!/usr/bin/env python
"""Text summary of an .ncu-rep (raw metrics + per-line hot spots) for profiles/:
   python tools/ncu_summary.py report.ncu-rep [B] > profiles/xxx.txt"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
skip = ["--launch-skip", sys.argv[3], "--launch-count", "1"] if len(sys.argv) > 3 else []  # which launch of the report
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + skip, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[2]
print("kernel:", v[h.index("Kernel Name")] if "Kernel Name" in h else "?")
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum"]
want += [k for k in h if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and "not_issued" not in k]
for w in want:
    if w in h:
        print(f"{w} = {v[h.index(w)]} {rows[1][h.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + skip, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[2]; samp = hdr.index('Warp Stall Sampling (All Samples)'); inst = hdr.index('Instructions Executed')
cur = None; lines = collections.defaultdict(lambda: [0, 0, ''])
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < len(hdr) or r[2] != '-' or not r[0].isdigit(): continue
    k = (cur, int(r[0])); lines[k][0] += float(r[samp] or 0); lines[k][1] += float(r[inst] or 0); lines[k][2] = r[1]
ts = sum(x[0] for x in lines.values()); ti = sum(x[1] for x in lines.values())
print(f"\nwarp instructions per instance: {ti / B:.0f} (B = {B}); stall samples {ts:.0f}")
# function-level roll-up by enclosing function (nearest preceding line that looks like a definition)
import os, re
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cheeta-mpc_b200", "csrc")
def funcs(fn):
    out = []
    try:
        for i, l in enumerate(open(os.path.join(root, fn)).read().split("\n")):
            m = re.match(r"^(?:template.*)?\s*(?:static\s+)?(?:__device__|__global__|__host__)[^;{]*?\b(\w+)\s*\(", l)
            if m: out.append((i + 1, m.group(1)))
    except OSError:
        pass
    return out
agg = collections.defaultdict(lambda: [0, 0])
for (f, ln), x in lines.items():
    fs = [n for (l, n) in funcs(f) if l <= ln] if f.endswith((".cuh", ".cu")) else []
    agg[(f, fs[-1] if fs else "-")][0] += x[0]; agg[(f, fs[-1] if fs else "-")][1] += x[1]
print("\nby function (source attribution of inlined code):")
for (f, n), x in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if x[1] / ti >= 0.004:
        print(f"  {f[:18]:18s} {n:26s} insts {100*x[1]/ti:5.1f}% ({x[1]/B:6.0f}/instance)  samples {100*x[0]/ts:5.1f}%")
print("\nhottest source lines:")
for (f, ln), x in sorted(lines.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"  {f[:16]:16s}{ln:5d} samp {100*x[0]/ts:5.1f}% inst {100*x[1]/ti:5.1f}%  {x[2].strip()[:100]}")
