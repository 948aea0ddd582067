#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/ncu_case.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, __graft_entry__ as ge
pkg = ge.load_package(); wl = pkg.workloads
cfg = wl.default_config(10); st, ds, di = wl.make_batch(cfg, 296)
m = pkg.CentroidalMPC.from_dict(cfg); m.SetupMPC(296)
for _ in range(3):
    out = m.UpdateMPCBatch(st, ds, di, want_lam=False)
print(out['stats'])
PY
python /tmp/ncu_case.py > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cmpc_solve_kernel -s 2 -c 1 -o gpurun_out/prof_v1 python /tmp/ncu_case.py > gpurun_out/ncu_v1.log 2>&1
tail -3 gpurun_out/ncu_plain.log gpurun_out/ncu_v1.log
