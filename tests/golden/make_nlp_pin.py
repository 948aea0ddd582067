"""Model-fidelity pin: the REFERENCE'S OWN non-convex NLP (CentroidalMPC.cpp:102-276) restated
for SciPy and solved on the reference driver's fixture and its variants, next to what this repo computes.

Restatement (single shooting over the same variables the reference declares at :109-134):
  variables   contact forces of the stance leg-steps (swing forces are pinned by 0 <= F f <= 0, :199), foot velocities of the
              swing leg-steps (a stance foot does not move: (1 - contact) * foot_vel, :94);
  dynamics    explicit Euler with the TRUE lever arm foot_pos - com_pos, both trajectory variables (:85-92);
  cost        :203-232 verbatim (omega_k inside the square, weight index map by code, foot-position tracking over all
              N + 1 nodes, force tracking against m g / #stance, force-rate via diff);
  constraints friction pyramid and force limits (:179-200), step box on foot_pos - des_foot_pos for nodes 1..N (:198).
Gradients come from torch autograd (float64); the solver is scipy.optimize SLSQP started at this repo's QP solution
(the NLP is non-convex; IPOPT in the reference starts at zero -- the local optimum next to the convex model is the
one a convexification can be compared with).

Stored in tests/golden/nlp_pin_v1.npz per case: the NLP's forces and foot positions, the frozen-arm QP's forces
(oracle), the forces at the fixed point of the arm re-linearisation (what cmpc_solve_batch_sqp converges to, emulated
here with the oracle), and the two gaps  max|f - f_nlp| / max|f_nlp|.
Run from the repo root:  python tests/golden/make_nlp_pin.py
"""
import os
import sys

import numpy as np
import torch
from scipy.optimize import minimize

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import numpy_mirror as nm  # noqa: E402

GRAV = 9.81
STEP_LB = np.array([-0.2, -0.2, -0.1])  # CentroidalMPC.cpp:30
STEP_UB = np.array([0.2, 0.2, 0.1])     # CentroidalMPC.cpp:31


class ReferenceNLP:
    def __init__(self, cfg, state, des_state, des_inputs):
        self.cfg = cfg
        self.N, self.L, self.m, self.dt = cfg["horizon"], cfg["num_legs"], cfg["mass"], cfg["dt"]
        self.w = np.asarray(cfg["weights"], float)
        x0, feet, dc, dv, dl, contact, dfoot = nm.unpack(cfg, state, des_state, des_inputs)
        self.x0, self.feet, self.dc, self.dv, self.dl, self.contact, self.dfoot = x0, feet, dc, dv, dl, contact, dfoot
        N, L = self.N, self.L
        self.stance = [(j, i) for j in range(N) for i in range(L) if contact[i, j] > 0]
        self.swing = [(j, i) for j in range(N) for i in range(L) if not contact[i, j] > 0]
        self.nf, self.nv = 3 * len(self.stance), 3 * len(self.swing)
        colsum = contact.sum(axis=0)
        self.fdes = np.zeros((N, L, 3))
        for (j, i) in self.stance:
            self.fdes[j, i, 2] = self.m * GRAV / colsum[j]          # :331-333
        self.t = lambda a: torch.tensor(a, dtype=torch.float64)

    def unpack(self, x):
        """x -> forces [N, L, 3] (zeros on swing legs), foot velocities [N, L, 3] (zeros on stance legs); torch."""
        N, L = self.N, self.L
        f = torch.zeros(N, L, 3, dtype=torch.float64)
        v = torch.zeros(N, L, 3, dtype=torch.float64)
        if self.nf:
            js, is_ = zip(*self.stance)
            f = f.index_put((torch.tensor(js), torch.tensor(is_)), x[:self.nf].reshape(-1, 3))
        if self.nv:
            js, is_ = zip(*self.swing)
            v = v.index_put((torch.tensor(js), torch.tensor(is_)), x[self.nf:].reshape(-1, 3))
        return f, v

    def rollout(self, x):
        N, L, dt, m = self.N, self.L, self.dt, self.m
        f, v = self.unpack(x)
        c, vel, am = [self.t(self.x0[0:3])], [self.t(self.x0[3:6])], [self.t(self.x0[6:9])]
        p = [self.t(self.feet)]                                           # [L, 3] per node
        ce = self.t(np.maximum(self.contact, 0.0))                       # [L, N]
        g = self.t([0.0, 0.0, -GRAV])
        for k in range(N):
            acc = g + (ce[:, k, None] / m * f[k]).sum(0)                 # :85
            ld = (ce[:, k, None] * torch.linalg.cross(p[k] - c[k][None, :], f[k])).sum(0)   # :86
            c.append(c[k] + vel[k] * dt); vel.append(vel[k] + acc * dt); am.append(am[k] + ld * dt)   # :90-92
            p.append(p[k] + (1.0 - ce[:, k, None]) * v[k] * dt)          # :94
        return torch.stack(c), torch.stack(vel), torch.stack(am), torch.stack(p), f

    def cost(self, x):
        N, L, w = self.N, self.L, self.w
        c, vel, am, p, f = self.rollout(x)
        om = self.t([(w[2] / 2) * np.exp(-k) + w[2] / 2 for k in range(N + 1)])           # :203-206
        dc, dv, dl = self.t(self.dc), self.t(self.dv), self.t(self.dl)
        J = w[0] * ((c[:, 0] - dc[:, 0]) ** 2).sum() + w[1] * ((c[:, 1] - dc[:, 1]) ** 2).sum() + ((om * (c[:, 2] - dc[:, 2])) ** 2).sum()
        for a in range(3):
            J = J + w[3 + a] * ((vel[:, a] - dv[:, a]) ** 2).sum() + w[6 + a] * ((am[:, a] - dl[:, a]) ** 2).sum()
        dfoot, fdes = self.t(self.dfoot), self.t(self.fdes)
        for i in range(L):
            for a in range(3):
                J = J + w[9 + 3 * i + a] * ((p[:, i, a] - dfoot[i, :, a]) ** 2).sum()                    # :219-221
                J = J + w[9 + 3 * L + 3 * i + a] * ((f[:, i, a] - fdes[:, i, a]) ** 2).sum()             # :223-225
                J = J + w[9 + 6 * L + 3 * i + a] * ((f[1:, i, a] - f[:-1, i, a]) ** 2).sum()             # :227-231
        return J

    def fun(self, xnp):
        x = torch.tensor(xnp, dtype=torch.float64, requires_grad=True)
        J = self.cost(x)
        J.backward()
        return float(J), x.grad.numpy().copy()

    def linear_constraints(self):
        """G x >= h: pyramid rows both sides (:186-199), step box on foot_pos - des_foot_pos, nodes 1..N (:198)."""
        N, L, dt = self.N, self.L, self.dt
        n = self.nf + self.nv
        G, h = [], []
        for s, (j, i) in enumerate(self.stance):
            mu = self.cfg["mu"][i]
            F = np.array([[-1, 0, mu], [1, 0, mu], [0, -1, mu], [0, 1, mu], [0, 0, 1.0]])
            ub = self.contact[i, j] * np.array([5000.0] * 4 + [self.m * GRAV * L])
            for r in range(5):
                row = np.zeros(n); row[3 * s:3 * s + 3] = F[r]
                G.append(row); h.append(0.0)
                G.append(-row); h.append(-ub[r])
        sw = {ji: s for s, ji in enumerate(self.swing)}
        for i in range(L):
            for k in range(1, N + 1):
                for a in range(3):
                    row = np.zeros(n)
                    for j in range(k):
                        if (j, i) in sw:
                            row[self.nf + 3 * sw[(j, i)] + a] = dt
                    base = self.feet[i, a] - self.dfoot[i, k, a]
                    G.append(row.copy()); h.append(STEP_LB[a] - base)
                    G.append(-row); h.append(base - STEP_UB[a])
        return np.array(G), np.array(h)

    def solve(self, f_init):
        x0 = np.zeros(self.nf + self.nv)
        for s, (j, i) in enumerate(self.stance):
            x0[3 * s:3 * s + 3] = f_init[j, i]
        G, h = self.linear_constraints()
        keep = np.abs(G).sum(axis=1) > 0          # rows without variables (stance-from-start feet) are data, not constraints
        assert (h[~keep] <= 1e-12).all(), "the fixture's own foot data violate the step box"
        G, h = G[keep], h[keep]
        J0, _ = self.fun(x0)
        sc = 1.0 / max(1.0, abs(J0))
        r = minimize(lambda x: tuple(sc * y for y in self.fun(x)), x0, jac=True, method="SLSQP",
                     constraints=[dict(type="ineq", fun=lambda x: G @ x - h, jac=lambda x: G)],
                     options=dict(ftol=1e-16, maxiter=3000))
        with torch.no_grad():
            c, vel, am, p, f = self.rollout(torch.tensor(r.x, dtype=torch.float64))
        return dict(forces=f.numpy(), foot_pos=p.numpy(), com=c.numpy(), cost=float(self.fun(r.x)[0]), cost_start=J0,
                    nit=r.nit, msg=r.message, max_violation=float(np.maximum(h - G @ r.x, 0).max()))


def nonlinear_com_path(cfg, state, contact, dfoot_true, F):
    """COM path of forces F [N, L, 3] through the reference plant (:85-92) with the true lever arms foot - com,
    feet following the desired foot positions (the QP's view of the feet)."""
    N, L, dt, m = cfg["horizon"], cfg["num_legs"], cfg["dt"], cfg["mass"]
    c, v = state[0:3].copy(), state[3:6].copy()
    path = [c.copy()]
    for k in range(N):
        acc = np.array([0, 0, -GRAV]) + sum(max(contact[i, k], 0) / m * F[k, i] for i in range(L))
        c = c + v * dt; v = v + acc * dt
        path.append(c.copy())
    return np.array(path)


def relinearised_fixed_point(pkg, orc, cfg, st, ds, di, iters=8):
    """What cmpc_solve_batch_sqp iterates (cmpc_api.cu relinearize_kernel): arms = des_foot_pos - c(j), c the COM path the
    previous forces produce through the nonlinear plant; implemented by shifting the desired foot positions."""
    N, L = cfg["horizon"], cfg["num_legs"]
    cc = pkg.make_config(cfg)
    x0, feet, dc, dv, dl, contact, dfoot = nm.unpack(cfg, st, ds, di)
    di_lin = di.copy()
    out = []
    for it in range(iters + 1):
        res = orc.solve_batch(cc, st[None], ds[None], di_lin[None])
        F = res["forces"][0].reshape(L, N, 3).transpose(1, 0, 2)
        out.append(F)
        path = nonlinear_com_path(cfg, st, contact, dfoot, F)
        D = di_lin.reshape(L, 4 * N + 3)
        D0 = di.reshape(L, 4 * N + 3)
        for i in range(L):
            for j in range(N):
                D[i, N + 3 * j:N + 3 * j + 3] = D0[i, N + 3 * j:N + 3 * j + 3] + (dc[j] - path[j])
    return out


def main():
    pkg, orc = ge.load_package(), ge.load_oracle()
    wl = pkg.workloads
    blob = {}
    cases = [(n,) + getattr(wl, n)() for n in ("fixture_f1", "fixture_f1_intended", "fixture_f1_n10")]
    hard = wl.hard_config(10, 0.3)
    st, ds, di = wl.make_batch(hard, 2, gaits=("trot", "bound"))
    cases += [("hard_trot_n10", hard, st[0], ds[0], di[0]), ("hard_bound_n10", hard, st[1], ds[1], di[1])]
    for name, cfg, st, ds, di in cases:
        N, L = cfg["horizon"], cfg["num_legs"]
        seq = relinearised_fixed_point(pkg, orc, cfg, st, ds, di)
        nlp = ReferenceNLP(cfg, st, ds, di)
        sol = nlp.solve(seq[-1])
        fn = sol["forces"]
        sc = np.abs(fn).max()
        gap_qp, gap_fp = np.abs(seq[0] - fn).max() / sc, np.abs(seq[-1] - fn).max() / sc
        foot_shift = np.abs(sol["foot_pos"] - np.transpose(nlp.dfoot, (1, 0, 2))).max()
        print(f"{name}: NLP cost {sol['cost']:.9g} (start {sol['cost_start']:.9g}), {sol['nit']} its, violation {sol['max_violation']:.1e}; "
              f"gap frozen-arm QP {gap_qp:.3e}, gap re-linearised fixed point {gap_fp:.3e}, "
              f"fixed-point increments {[float(np.abs(seq[k + 1] - seq[k]).max() / sc) for k in range(4)]}, max foot shift {foot_shift:.2e}", flush=True)
        pre = name + "/"
        blob[pre + "cfg"] = np.array([cfg["mass"], L, N, cfg["dt"], cfg["disc_mode"]] + list(cfg["mu"]) + list(cfg["weights"]))
        blob[pre + "state"], blob[pre + "des_state"], blob[pre + "des_inputs"] = st, ds, di
        blob[pre + "forces_nlp"], blob[pre + "foot_pos_nlp"] = fn, sol["foot_pos"]
        blob[pre + "forces_qp"], blob[pre + "forces_fixed_point"] = seq[0], seq[-1]
        blob[pre + "forces_sqp2"] = seq[2]
        blob[pre + "gaps"] = np.array([gap_qp, gap_fp])
        blob[pre + "cost_nlp"] = np.array([sol["cost"], sol["cost_start"]])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nlp_pin_v1.npz"), **blob)


if __name__ == "__main__":
    main()
