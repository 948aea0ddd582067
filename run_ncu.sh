#!/bin/bash
# ncu capture of the solve kernel on one headline-config batch (B=4096), source-level
mkdir -p gpurun_out
cat > /tmp/ncu_case.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, __graft_entry__ as ge
pkg = ge.load_package(); wl = pkg.workloads
B = 4096
cfg = wl.default_config(10); st, ds, di = wl.make_batch(cfg, B)
m = pkg.CentroidalMPC.from_dict(cfg); m.SetupMPC(B)
for _ in range(3):
    out = m.UpdateMPCBatch(st, ds, di, want_lam=False)
print(out['stats'])
PY
python /tmp/ncu_case.py > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cmpc_solve_kernel -s 2 -c 1 -f -o gpurun_out/prof python /tmp/ncu_case.py > gpurun_out/ncu.log 2>&1
tail -n 3 gpurun_out/ncu_plain.log gpurun_out/ncu.log
