"""Small batches through every kernel family (for compute-sanitizer): dense presolve W=1/4/8, Riccati,
interior-point kernel with active rows, closed loop with warm start."""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, __graft_entry__ as ge
from conftest import hard_config
pkg = ge.load_package(); wl = pkg.workloads
def run(tag, cfg, B, gaits):
    st, ds, di = wl.make_batch(cfg, B, gaits=gaits)
    m = pkg.CentroidalMPC.from_dict(cfg); m.SetupMPC(B)
    out = m.UpdateMPCBatch(st, ds, di)
    print(tag, np.bincount(out["status"], minlength=5), "iters", out["iters"].mean(), "kkt", out["kkt"].max())
    m.close()
run("N=10 mixed", wl.default_config(10), 96, wl.GAITS)
run("N=6", wl.default_config(6), 40, wl.GAITS)
run("N=16 stand (W=8 dense)", dict(wl.default_config(16), qp_backend=1), 24, ("stand",))
run("N=30 mixed (Riccati)", wl.default_config(30), 40, wl.GAITS)
run("N=10 Riccati everywhere", dict(wl.default_config(10), qp_backend=2), 64, wl.GAITS)
run("hard mu 0.3", hard_config(wl, 10, 0.3), 64, wl.GAITS)
cfg = hard_config(wl, 10, 0.3)
st, ds, di = wl.make_batch(cfg, 48, gaits=wl.GAITS)
m = pkg.CentroidalMPC.from_dict(cfg); m.SetupMPC(48)
r = m.Rollout(st, ds, di, 6, warm_start=1)
print("rollout warm", np.bincount(r["status_or"]))
m.close()
