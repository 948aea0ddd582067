#!/bin/bash
# launch list + full capture of the phase-split condensed interior-point kernels on the tracking-heavy trot workload
mkdir -p gpurun_out
cat > /tmp/ncu_split.py <<'PY'
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch, __graft_entry__ as ge
pkg = ge.load_package(); wl = pkg.workloads
B = 4096
cfg = dict(wl.hard_config(10, 0.3), qp_backend=1); st, ds, di = wl.make_batch(cfg, B, gaits=("trot",))
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
python /tmp/ncu_split.py > gpurun_out/ncu_split_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/split_launches.csv python /tmp/ncu_split.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:cmpc_solve_kernel -s 3 -c 3 -f -o gpurun_out/prof_split python /tmp/ncu_split.py > gpurun_out/ncu_split.log 2>&1
tail -n 1 gpurun_out/ncu_split_plain.log; grep -E "cmpc_|stats" gpurun_out/split_launches.csv | tail -12 | cut -d, -f5,12-
