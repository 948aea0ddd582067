"""Pins the oracle with arithmetic this repo did not write: every golden case (tests/golden/golden_v1.npz inputs,
incl. the ones with active friction rows) is solved again by SciPy's SLSQP (Kraft's sequential least-squares QP
code) on the swing-eliminated QP   min 1/2 u'Hu + g'u,  0 <= C u <= ub   built by the NumPy mirror, and, for a
subset, by SciPy's trust-constr interior-point method.  SLSQP stops at its own accuracy (1e-8 .. 2e-6 relative on
these problems; restarts do not move it), so its ACTIVE SET is then refined by one LAPACK solve of the equality-
constrained KKT system (scipy.linalg QR for the row rank, numpy.linalg.solve) -- still no arithmetic of this repo.
The script asserts, before writing, that SciPy and the C oracle agree (raw SLSQP forces <= 5e-6 relative per
instance, refined <= 1e-9, identical active set) and stores SciPy's solutions in
tests/golden/scipy_pin_v1.npz; tests/test_oracle.py::test_scipy_pin and the GPU parity tests compare against the
stored vectors, so the pin travels without SciPy having to run on the GPU box.
Run from the repo root:  python tests/golden/make_scipy_pin.py        (about two minutes)
"""
import os
import sys
import time

import numpy as np
import scipy
import scipy.linalg as sla
from scipy.optimize import Bounds, LinearConstraint, minimize

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import numpy_mirror as nm  # noqa: E402
from conftest import to_step_major  # noqa: E402

ACT_TOL = 1e-7   # a row counts as active when its slack is below ACT_TOL * (1 + max|u|)


def reduced_qp(cfg, st, ds, di):
    qp = nm.build_qp(cfg, st, ds, di)
    free = ~qp["pinned"]
    C, lb, ub, tags = nm.constraints(cfg, qp["contact"])
    return qp["H"][np.ix_(free, free)], qp["g"][free], C[:, free], ub, tags, free, qp


def feasible_start(cfg, qp, free, C, ub, tags):
    """f = (0, 0, min(fz_ref, fz_max / 2, 5000 / (2 mu))) per stance leg-step: strictly inside every row."""
    N, L = cfg["horizon"], cfg["num_legs"]
    u = np.zeros(3 * L * N)
    for (j, i, r) in tags:
        if r == 4:
            fz = qp["Uref"][3 * L * j + 3 * i + 2]
            u[3 * L * j + 3 * i + 2] = min(fz, 0.5 * ub[tags.index((j, i, 4))], 0.5 * ub[tags.index((j, i, 0))] / cfg["mu"][i])
    return u[free]


def solve_slsqp(H, g, C, ub, u0):
    f = lambda u: (0.5 * u @ H @ u + g @ u, H @ u + g)
    cons = [dict(type="ineq", fun=lambda u: C @ u, jac=lambda u: C),
            dict(type="ineq", fun=lambda u: ub - C @ u, jac=lambda u: -C)]
    sc = 1.0 / max(1.0, np.abs(g).max())   # SLSQP's ftol is absolute: scale the objective
    r = minimize(lambda u: tuple(sc * x for x in f(u)), u0, jac=True, method="SLSQP", constraints=cons,
                 options=dict(ftol=1e-16, maxiter=2000))
    return r.x, r


def solve_trust_constr(H, g, C, ub, u0):
    r = minimize(lambda u: 0.5 * u @ H @ u + g @ u, u0, jac=lambda u: H @ u + g, hess=lambda u: H, method="trust-constr",
                 constraints=[LinearConstraint(C, np.zeros(len(ub)), ub)],
                 options=dict(gtol=1e-12, xtol=1e-14, barrier_tol=1e-12, maxiter=3000))
    return r.x, r


def refine(H, g, C, ub, u):
    """Equality-constrained KKT solve on the rows SLSQP left active (independent rows picked by pivoted QR)."""
    y = C @ u
    tol = ACT_TOL * (1 + np.abs(u).max())
    lo, hi = y <= tol, ub - y <= tol
    A = np.vstack([C[lo], C[hi]]); rhs = np.concatenate([np.zeros(lo.sum()), ub[hi]])
    if len(A) == 0:
        return np.linalg.solve(H, -g)
    _, R, piv = sla.qr(A.T, mode="economic", pivoting=True)
    rk = int((np.abs(np.diag(R)) > 1e-10 * abs(R[0, 0])).sum())
    A, rhs = A[piv[:rk]], rhs[piv[:rk]]
    K = np.block([[H, A.T], [A, np.zeros((rk, rk))]])
    return np.linalg.solve(K, np.concatenate([-g, rhs]))[:len(u)]


def active_set(C, ub, u, tags, N, L):
    y = C @ u
    tol = ACT_TOL * (1 + np.abs(u).max())
    a = np.zeros((N, L), np.uint16)
    for t, (j, i, r) in enumerate(tags):
        if y[t] <= tol:
            a[j, i] |= 1 << r
        if ub[t] - y[t] <= tol:
            a[j, i] |= 1 << (5 + r)
    return a


def main():
    pkg, orc = ge.load_package(), ge.load_oracle()
    z = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
    blob = {"scipy_version": np.array(scipy.__version__)}
    worst = 0.0
    for name in sorted({k.split("/")[0] for k in z.files}):
        v = z[name + "/cfg"]
        L = int(v[1]); N = int(v[2])
        cfg = dict(mass=float(v[0]), num_legs=L, horizon=N, dt=float(v[3]), disc_mode=int(v[4]),
                   mu=list(v[5:5 + L]), weights=list(v[5 + L:5 + L + 9 + 9 * L]))
        st, ds, di = z[name + "/state"], z[name + "/des_state"], z[name + "/des_inputs"]
        ref = orc.solve_batch(pkg.make_config(cfg), st, ds, di)
        Uo = to_step_major(ref["forces"], N, L)
        Us, As, Ts, Rs = [], [], [], []
        for b in range(len(st)):
            t0 = time.time()
            H, g, C, ub, tags, free, qp = reduced_qp(cfg, st[b], ds[b], di[b])
            u0 = feasible_start(cfg, qp, free, C, ub, tags)
            u, r = solve_slsqp(H, g, C, ub, u0)
            U = np.zeros(3 * L * N); U[free] = u
            err = np.abs(U - Uo[b]).max() / np.abs(Uo[b]).max()
            act = active_set(C, ub, u, tags, N, L)
            oact = ref["active"][b] & 0x3FF
            assert err <= 5e-6, (name, b, err, r.message)
            assert np.array_equal(act, oact), (name, b, int((act != oact).sum()))
            Ur = np.zeros(3 * L * N); Ur[free] = refine(H, g, C, ub, u)
            rerr = np.abs(Ur - Uo[b]).max() / np.abs(Uo[b]).max()
            assert rerr <= 1e-9, (name, b, rerr)
            worst = max(worst, err)
            tc_err = np.nan
            if b == 0 and N <= 10:   # second opinion from an interior-point code
                u2, r2 = solve_trust_constr(H, g, C, ub, u0)
                tc_err = np.abs(u2 - u).max() / np.abs(u).max()
                assert tc_err <= 5e-6, (name, b, tc_err)
            Us.append(U); As.append(act); Ts.append(tc_err); Rs.append(Ur)
            print(f"{name}[{b}] n={len(u)} rows={len(ub)} active={int(sum(bin(int(x)).count('1') for x in act.ravel()))} "
                  f"slsqp-vs-oracle {err:.1e} refined {rerr:.1e} trust-constr-vs-slsqp {tc_err:.1e} ({time.time() - t0:.1f}s, {r.nit} its)", flush=True)
        blob[name + "/U_slsqp"] = np.array(Us)
        blob[name + "/U_refined"] = np.array(Rs)
        blob[name + "/active_slsqp"] = np.array(As)
        blob[name + "/trust_constr_err"] = np.array(Ts)
    print("worst SLSQP-vs-oracle relative force error", worst)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "scipy_pin_v1.npz"), **blob)


if __name__ == "__main__":
    main()
