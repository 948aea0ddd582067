"""Randomised configurations (horizon, weights over six decades, friction, time step, discretisation, back-end knobs)
against the CPU oracle: forces, status, active set, iteration counts.  python tools/random_config_sweep.py [n_configs] [seed]"""
import sys; sys.path.insert(0, '.')
import numpy as np, __graft_entry__ as ge
pkg, orc = ge.load_package(), ge.load_oracle(); wl = pkg.workloads
ncfg = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for t in range(ncfg):
    N = int(rng.choice([3, 5, 8, 10, 12, 16, 20, 30]))
    w = np.concatenate([10 ** rng.uniform(0, 5, 9), 10 ** rng.uniform(-2, 1, 12), 10 ** rng.uniform(-3, 0, 12), 10 ** rng.uniform(-4, -1, 12)])
    cfg = wl.default_config(N, dt=float(rng.uniform(0.005, 0.05)), mu=[float(rng.uniform(0.05, 1.0))] * 4, weights=w, disc_mode=int(rng.integers(0, 2)))
    cfg = dict(cfg, qp_backend=int(rng.choice([0, 0, 1, 2])), presolve=int(rng.choice([1, 1, 0])))
    if cfg["qp_backend"] == 1 and N > 16: cfg["qp_backend"] = 0
    B = 96 if N <= 16 else 32
    st, ds, di = wl.make_batch(cfg, B, first=1000 * t, gaits=wl.GAITS, hard_fraction=float(rng.choice([0.25, 0.5, 1.0])))
    m = pkg.CentroidalMPC.from_dict(cfg); m.SetupMPC(B)
    out = m.UpdateMPCBatch(st, ds, di); m.close()
    ref = orc.solve_batch(m.cfg, st, ds, di, nthreads=16)
    ok = ref["status"] == 0
    scale = np.abs(ref["forces"]).max(axis=1, keepdims=True) + 1e-300
    err = float((np.abs(out["forces"] - ref["forces"]) / scale)[ok].max()) if ok.any() else 0.0
    ds_ = int((out["status"] != ref["status"]).sum()); da = int((out["active"] != ref["active"]).any(axis=(1, 2))[ok].sum())
    di_ = int(np.abs(out["iters"] - ref["iters"]).max())
    flag = "" if (err <= 1e-6 and ds_ == 0 and da == 0) else "   <-- MISMATCH"
    bad += bool(flag)
    print(f"cfg {t:3d} N={N:2d} dt={cfg['dt']:.3f} mu={cfg['mu'][0]:.2f} zoh={cfg['disc_mode']} backend={cfg['qp_backend']} presolve={cfg['presolve']}: "
          f"err {err:.1e} status!= {ds_} active!= {da} |iters diff| {di_} settled {100 * (out['iters'] == 0).mean():3.0f}% status {np.bincount(out['status'], minlength=5).tolist()} "
          f"ref {np.bincount(ref['status'], minlength=5).tolist()}{flag}", flush=True)
print("ALL OK" if bad == 0 else f"{bad} configurations with mismatches")
