#!/bin/bash
# Round-2 profile set: launch list of the bench command, then full captures of the kernels that matter:
#   the presolve kernel (headline), the two phase kernels of the condensed interior point (constrained, horizon 10),
#   the stage-wise interior-point kernel (constrained, horizon 30).  Each command runs once without ncu first.
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
cat > /tmp/ncu_case.py <<'PY'
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch, __graft_entry__ as ge
pkg = ge.load_package(); wl = pkg.workloads
case = sys.argv[1]
if case == "headline":
    cfg = wl.default_config(10); B = 4096; gaits = ("trot",)
elif case == "hard10":
    cfg = wl.hard_config(10, 0.3); B = 4096; gaits = ("trot",)
else:
    cfg = wl.hard_config(30, 0.3); B = 2048; gaits = wl.GAITS
st, ds, di = wl.make_batch(cfg, B, gaits=gaits)
dev = torch.device('cuda', 0)
m = pkg.CentroidalMPC.from_dict(cfg); m.SetupMPC(B)
d = [torch.from_numpy(a).to(dev) for a in (st, ds, di)]
f = torch.zeros(B, m.n_forces, dtype=torch.float64, device=dev); s = torch.zeros(B, dtype=torch.int32, device=dev)
it = torch.zeros(B, dtype=torch.int32, device=dev); kk = torch.zeros(B, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
for _ in range(3):
    stats = pkg.CmpcStats()
    m.solve_device(B, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), f.data_ptr(), s.data_ptr(), it.data_ptr(), kk.data_ptr(), stats=stats)
print(case, stats.as_dict())
PY
for c in headline hard10 hard30; do
  python /tmp/ncu_case.py $c > gpurun_out/r02_${c}_plain.log 2>&1 || exit 1
done
# third call of each case (warm): every cmpc kernel of that call
ncu --set full --clock-control none --import-source on -k regex:cmpc_presolve_kernel -s 2 -c 1 -f -o gpurun_out/r02_prof_presolve python /tmp/ncu_case.py headline > gpurun_out/r02_ncu_headline.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cmpc_solve_kernel -s 4 -c 2 -f -o gpurun_out/r02_prof_dense_phases python /tmp/ncu_case.py hard10 > gpurun_out/r02_ncu_hard10.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cmpc_ripm_kernel -s 2 -c 1 -f -o gpurun_out/r02_prof_ripm_n30 python /tmp/ncu_case.py hard30 > gpurun_out/r02_ncu_hard30.log 2>&1
tail -n 1 gpurun_out/r02_*_plain.log
