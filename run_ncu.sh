#!/bin/bash
# ncu capture of one kernel on one headline-config batch (B=4096, device-resident), source-level.
# KERN=cmpc_presolve_kernel (default, the headline's dominant kernel) or cmpc_solve_kernel with PRESOLVE=0; OUT=report name
KERN=${KERN:-cmpc_presolve_kernel}; OUT=${OUT:-prof}; export PRESOLVE=${PRESOLVE:-1}
mkdir -p gpurun_out
cat > /tmp/ncu_case.py <<'PY'
import sys; sys.path.insert(0, '.')
import os, numpy as np, torch, __graft_entry__ as ge
pkg = ge.load_package(); wl = pkg.workloads
B = 4096
cfg = dict(wl.default_config(10), presolve=int(os.environ.get('PRESOLVE', '1'))); st, ds, di = wl.make_batch(cfg, B)
dev = torch.device('cuda', 0)
m = pkg.CentroidalMPC.from_dict(cfg); m.SetupMPC(B)
d = [torch.from_numpy(a).to(dev) for a in (st, ds, di)]
f = torch.zeros(B, 120, dtype=torch.float64, device=dev); s = torch.zeros(B, dtype=torch.int32, device=dev)
it = torch.zeros(B, dtype=torch.int32, device=dev); kk = torch.zeros(B, dtype=torch.float64, device=dev)
for _ in range(3):
    stats = pkg.CmpcStats()
    m.solve_device(B, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), f.data_ptr(), s.data_ptr(), it.data_ptr(), kk.data_ptr(), stats=stats)
print(stats.as_dict())
PY
python /tmp/ncu_case.py > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$KERN -s 4 -c 1 -f -o gpurun_out/$OUT python /tmp/ncu_case.py > gpurun_out/ncu.log 2>&1
tail -n 3 gpurun_out/ncu_plain.log gpurun_out/ncu.log
