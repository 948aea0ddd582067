#!/bin/bash
# ncu capture of the Riccati presolve kernel on a horizon-30 stand batch (n = 360), source-level
mkdir -p gpurun_out
cat > /tmp/ncu_ric.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch, __graft_entry__ as ge
pkg = ge.load_package(); wl = pkg.workloads
B = 1024
cfg = wl.default_config(30); st, ds, di = wl.make_batch(cfg, B, gaits=("stand",))
dev = torch.device('cuda', 0)
m = pkg.CentroidalMPC.from_dict(cfg); m.SetupMPC(B)
d = [torch.from_numpy(a).to(dev) for a in (st, ds, di)]
f = torch.zeros(B, m.n_forces, dtype=torch.float64, device=dev); s = torch.zeros(B, dtype=torch.int32, device=dev)
it = torch.zeros(B, dtype=torch.int32, device=dev); kk = torch.zeros(B, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
for _ in range(3):
    stats = pkg.CmpcStats()
    m.solve_device(B, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), f.data_ptr(), s.data_ptr(), it.data_ptr(), kk.data_ptr(), stats=stats)
print(stats.as_dict())
PY
python /tmp/ncu_ric.py > gpurun_out/ncu_ric_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cmpc_riccati_kernel -s 2 -c 1 -f -o gpurun_out/prof_riccati python /tmp/ncu_ric.py > gpurun_out/ncu_ric.log 2>&1
tail -n 2 gpurun_out/ncu_ric_plain.log gpurun_out/ncu_ric.log
