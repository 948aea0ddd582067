"""Generates tests/golden/golden_v1.npz.

The reference pins no outputs (UpdateMPC returns {}, CentoidMPCTest.cpp asserts nothing;
CasADi/IPOPT absent), so the golden vectors come from this repo's C oracle and are
cross-checked here, before being written, by the independent NumPy mirror:
  * H, g against the dense NumPy build (<= 1e-13 relative),
  * the solution against an independent KKT evaluation (<= 1e-12) and, where it
    converges, the textbook active-set solver (<= 1e-8 relative).
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import numpy_mirror as nm  # noqa: E402
from conftest import hard_config, to_step_major  # noqa: E402


def cases(wl):
    out = []
    for name in ("fixture_f1", "fixture_f1_intended", "fixture_f1_n10"):
        cfg, st, ds, di = getattr(wl, name)()
        out.append((name, cfg, st[None], ds[None], di[None]))
    cfg = wl.default_config(10)
    out.append(("config2_first8", cfg) + wl.make_batch(cfg, 8))
    out.append(("config4_first10", cfg) + wl.make_batch(cfg, 10, gaits=wl.GAITS))
    cfg = wl.default_config(10, disc_mode=1)
    out.append(("config2_zoh_first4", cfg) + wl.make_batch(cfg, 4, gaits=wl.GAITS))
    for mu in (0.8, 0.3, 0.1):
        cfg = hard_config(wl, 10, mu)
        out.append((f"hard_mu{mu}_first12", cfg) + wl.make_batch(cfg, 12, gaits=wl.GAITS))
    cfg = wl.default_config(30)
    out.append(("config3_first2", cfg) + wl.make_batch(cfg, 2, gaits=("trot", "stand")))
    cfg = hard_config(wl, 30, 0.3)
    out.append(("hard_n30_first2", cfg) + wl.make_batch(cfg, 2, gaits=("trot", "gallop")))
    return out


def main():
    pkg, orc = ge.load_package(), ge.load_oracle()
    wl = pkg.workloads
    blob = {}
    for name, cfg, st, ds, di in cases(wl):
        cc = pkg.make_config(cfg)
        res = orc.solve_batch(cc, st, ds, di)
        N, L = cfg["horizon"], cfg["num_legs"]
        U = to_step_major(res["forces"], N, L)
        Hs, gs = [], []
        for b in range(len(st)):
            H, g, s = orc.build_qp(cc, st[b], ds[b], di[b])
            qp = nm.build_qp(cfg, st[b], ds[b], di[b])
            assert np.abs(H - qp["H"]).max() <= 1e-13 * np.abs(qp["H"]).max(), name
            assert np.abs(g - qp["g"]).max() <= 1e-13 * max(1.0, np.abs(qp["g"]).max()), name
            C, lb, ub, tags = nm.constraints(cfg, qp["contact"])
            ll = np.array([res["lam"][b, 0, j, i, r] for (j, i, r) in tags])
            lu = np.array([res["lam"][b, 1, j, i, r] for (j, i, r) in tags])
            k = nm.kkt_residual(cfg, qp, U[b], ll, lu)
            assert res["status"][b] == 0 and k <= 1e-12, (name, b, res["status"][b], k)
            try:
                Ua, _, _, _ = nm.solve_active_set(cfg, qp)
                assert np.abs(Ua - U[b]).max() <= 1e-8 * np.abs(U[b]).max(), (name, b)
            except RuntimeError:
                pass  # textbook active-set cycles on degenerate vertices; KKT check above stands
            Hs.append(H); gs.append(g)
        pre = name + "/"
        blob[pre + "cfg"] = np.array([cfg["mass"], cfg["num_legs"], cfg["horizon"], cfg["dt"], cfg["disc_mode"]]
                                     + list(cfg["mu"]) + list(cfg["weights"]))
        blob[pre + "state"], blob[pre + "des_state"], blob[pre + "des_inputs"] = st, ds, di
        blob[pre + "forces"], blob[pre + "active"] = res["forces"], res["active"]
        blob[pre + "lam"], blob[pre + "iters"] = res["lam"], res["iters"]
        blob[pre + "H"] = np.array(Hs)   # full H at every horizon (the N = 30 matrices compress to a few hundred KB)
        blob[pre + "g"] = np.array(gs)
        blob[pre + "H_diag"] = np.array([np.diag(h) for h in Hs])
        blob[pre + "H_rowsum"] = np.array([h.sum(axis=1) for h in Hs])
        print(name, "ok", "iters", res["iters"].tolist())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), **blob)


if __name__ == "__main__":
    main()
