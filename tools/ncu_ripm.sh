#!/bin/bash
# ncu capture of the stage-wise interior-point kernel (cmpc_ripm.cu) on the tracking-heavy workload, source-level.
# usage: tools/ncu_ripm.sh <horizon> <batch> <gaits> <tag> <launches to skip> <launches to capture>
N=${1:-10}; B=${2:-2048}; G=${3:-stand}; TAG=${4:-ripm}; SKIP=${5:-2}; CNT=${6:-2}
mkdir -p gpurun_out
cat > /tmp/ncu_ripm.py <<PY
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch, __graft_entry__ as ge
from conftest import hard_config
pkg = ge.load_package(); wl = pkg.workloads
B = $B
cfg = dict(hard_config(wl, $N, 0.3), qp_backend=2); st, ds, di = wl.make_batch(cfg, B, gaits="$G".split(","))
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
python /tmp/ncu_ripm.py > gpurun_out/ncu_${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cmpc_ripm_kernel -s $SKIP -c $CNT -f -o gpurun_out/prof_${TAG} python /tmp/ncu_ripm.py > gpurun_out/ncu_${TAG}.log 2>&1
tail -n 2 gpurun_out/ncu_${TAG}_plain.log gpurun_out/ncu_${TAG}.log
