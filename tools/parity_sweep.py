"""Large randomized parity sweep GPU vs CPU oracle (beyond the pytest sizes): forces, status, active set, KKT, and the
route taken (iters == 0 <=> settled by the presolve) for easy / tracking-heavy workloads, both presolve back-ends."""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, __graft_entry__ as ge
from conftest import hard_config
pkg, orc = ge.load_package(), ge.load_oracle(); wl = pkg.workloads
def sweep(tag, cfg, B, gaits, first=0):
    st, ds, di = wl.make_batch(cfg, B, first=first, gaits=gaits)
    m = pkg.CentroidalMPC.from_dict(cfg); m.SetupMPC(B)
    out = m.UpdateMPCBatch(st, ds, di)
    ref = orc.solve_batch(m.cfg, st, ds, di, nthreads=16)
    scale = np.abs(ref["forces"]).max(axis=1, keepdims=True) + 1e-300
    err = (np.abs(out["forces"] - ref["forces"]) / scale).max()
    bad_status = int((out["status"] != ref["status"]).sum())
    bad_active = int((out["active"] != ref["active"]).any(axis=(1, 2)).sum())
    bad_route = int(((out["iters"] == 0) != (ref["iters"] == 0)).sum())
    bad_iters = int((out["iters"] != ref["iters"]).sum())
    ok = out["status"] <= 1
    print(f"{tag:38s} B={B:5d} max rel force err {err:.2e}  status!= {bad_status}  active!= {bad_active}  route!= {bad_route}  iters!= {bad_iters}"
          f"  max kkt {out['kkt'][ok].max():.1e}  settled {100*(out['iters']==0).mean():.0f}%  status {np.bincount(out['status'], minlength=5).tolist()}")
    m.close()
    return err <= 1e-6 and bad_status == 0 and bad_active == 0
allok = True
allok &= sweep("config2 trot N=10", wl.default_config(10), 8192, ("trot",), first=100000)
allok &= sweep("mixed gaits N=10", wl.default_config(10), 8192, wl.GAITS, first=200000)
allok &= sweep("mixed gaits N=10 ZOH", wl.default_config(10, disc_mode=1), 4096, wl.GAITS, first=300000)
for mu in (0.8, 0.3, 0.1):
    allok &= sweep(f"tracking-heavy mu={mu} N=10", hard_config(wl, 10, mu), 4096, wl.GAITS, first=400000)
allok &= sweep("tracking-heavy mu=0.8, Riccati all", dict(hard_config(wl, 10, 0.8), qp_backend=2), 2048, wl.GAITS, first=500000)
allok &= sweep("mixed gaits N=30", wl.default_config(30), 2048, wl.GAITS, first=600000)
allok &= sweep("tracking-heavy mu=0.3 N=30", hard_config(wl, 30, 0.3), 256, wl.GAITS, first=700000)
allok &= sweep("mixed gaits N=6", wl.default_config(6), 2048, wl.GAITS, first=800000)
print("ALL OK" if allok else "MISMATCH")
