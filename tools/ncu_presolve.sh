#!/bin/bash
# Full ncu capture of the presolve kernel on the headline workload (third, warm call).  $1 = tag of the report.
# The same command runs once without ncu first.
tag=${1:-presolve}
mkdir -p gpurun_out
cat > /tmp/ncu_case.py <<'PY'
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch, __graft_entry__ as ge
pkg = ge.load_package(); wl = pkg.workloads
cfg = wl.default_config(10); B = 4096
st, ds, di = wl.make_batch(cfg, B, gaits=("trot",))
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
python /tmp/ncu_case.py > gpurun_out/${tag}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:cmpc_presolve_kernel -s 2 -c 1 -f -o gpurun_out/${tag} python /tmp/ncu_case.py > gpurun_out/${tag}_ncu.log 2>&1
tail -n 1 gpurun_out/${tag}_plain.log
