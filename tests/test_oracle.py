"""CPU tests: the C oracle against the golden vectors, the independent NumPy mirror and
size-independent properties.  The reference pins nothing for this path (PARITY UNPINNED,
see oracle/cmpc_oracle.c header), so the oracle is pinned by independent arithmetic."""
import os

import numpy as np
import pytest

import numpy_mirror as nm
from conftest import hard_config, to_step_major

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz")


def golden_cases():
    z = np.load(GOLDEN)
    names = sorted({k.split("/")[0] for k in z.files})
    return z, names


def cfg_from_golden(z, name):
    v = z[name + "/cfg"]
    L = int(v[1])
    return dict(mass=float(v[0]), num_legs=L, horizon=int(v[2]), dt=float(v[3]), disc_mode=int(v[4]),
                mu=list(v[5:5 + L]), weights=list(v[5 + L:5 + L + 9 + 9 * L]))


@pytest.mark.parametrize("name", golden_cases()[1])
def test_oracle_reproduces_golden(pkg, orc, name):
    z, _ = golden_cases()
    cfg = cfg_from_golden(z, name)
    cc = pkg.make_config(cfg)
    st, ds, di = z[name + "/state"], z[name + "/des_state"], z[name + "/des_inputs"]
    res = orc.solve_batch(cc, st, ds, di)
    assert (res["status"] == 0).all()
    scale = np.abs(z[name + "/forces"]).max()
    assert np.abs(res["forces"] - z[name + "/forces"]).max() <= 1e-9 * scale
    assert np.array_equal(res["active"], z[name + "/active"])
    for b in range(len(st)):
        H, g, s = orc.build_qp(cc, st[b], ds[b], di[b])
        assert np.abs(g - z[name + "/g"][b]).max() <= 1e-13 * max(1.0, np.abs(g).max())
        assert np.abs(np.diag(H) - z[name + "/H_diag"][b]).max() <= 1e-13 * np.abs(H).max()
        assert np.abs(H.sum(axis=1) - z[name + "/H_rowsum"][b]).max() <= 1e-12 * np.abs(H).max()
        if name + "/H" in z.files:
            assert np.abs(H - z[name + "/H"][b]).max() <= 1e-13 * np.abs(H).max()


def test_fixture_f1_matches_reference_driver_values(wl):
    """F1 encodes CentoidMPCTest.cpp:12-107 under the memcpy semantics of CentroidalMPC.cpp:284-317."""
    cfg, st, ds, di = wl.fixture_f1()
    assert cfg["horizon"] == 6 and cfg["mass"] == 8 and cfg["dt"] == 0.01 and cfg["mu"] == [0.8] * 4
    assert len(st) == 21 and len(ds) == 63 and len(di) == 4 * 27
    assert np.count_nonzero(ds[54:]) == 0            # the 9 slots the comma initialiser never fills
    x0, feet, dc, dv, dl, contact, dfoot = nm.unpack(cfg, st, ds, di)
    # blocks are read at stride 3(N+1)=21, so des_com_pos node 6 is the driver's first "vel" row
    assert np.allclose(dc[6], [0.1, 0, 0]) and np.allclose(dc[0], [0.31, 0, 0.16])
    assert contact[:, :3].T.tolist() == [[1, 0, 1, 0]] * 3 and contact[:, 3:].T.tolist() == [[0, 1, 0, 1]] * 3
    assert np.allclose(dfoot[1][3], [0.43, -0.052, 0])
    # weight indexing follows the code (CentroidalMPC.cpp:219-231), not the driver's comments
    w = np.array(cfg["weights"])
    assert w[9:21].tolist() == [0.2, 0.2, 0.2, 0.3, 0.3, 0.3, 0.1, 0.1, 0.1, 0.2, 0.2, 0.2]


@pytest.mark.parametrize("disc", [0, 1])
@pytest.mark.parametrize("N", [1, 2, 6, 10])
def test_oracle_build_matches_numpy_mirror(pkg, orc, wl, N, disc):
    cfg = hard_config(wl, N, 0.4, disc_mode=disc)
    st, ds, di = wl.make_batch(cfg, 6, gaits=wl.GAITS)
    cc = pkg.make_config(cfg)
    for b in range(6):
        qp = nm.build_qp(cfg, st[b], ds[b], di[b])
        H, g, s = orc.build_qp(cc, st[b], ds[b], di[b])
        assert np.abs(H - qp["H"]).max() <= 1e-13 * np.abs(qp["H"]).max()
        assert np.abs(g - qp["g"]).max() <= 1e-13 * np.abs(qp["g"]).max()
        ev = np.linalg.eigvalsh(H)
        assert ev.min() > 0                           # K > 0  =>  strictly convex


def test_zoh_closed_form_equals_generic_c2d(wl):
    """B_zoh = B_euler + dt^2/2 [c/m I;0;0], d_zoh = d + [g dt^2/2;0;0]; the mirror
    obtains ZOH from the (terminating) exponential series of the augmented matrix."""
    cfg = wl.default_config(4, disc_mode=1)
    c = np.array([1.0, 0.0, 1.0, 1.0])
    r = np.random.default_rng(0).normal(size=(4, 3))
    A1, B1, d1 = nm.discretize(cfg, c, r)
    A0, B0, d0 = nm.discretize(dict(cfg, disc_mode=0), c, r)
    dt, m = cfg["dt"], cfg["mass"]
    assert np.array_equal(A0, A1)
    extra = np.zeros_like(B0)
    for i in range(4):
        extra[0:3, 3 * i:3 * i + 3] = 0.5 * dt * dt * c[i] / m * np.eye(3)
    assert np.allclose(B1, B0 + extra, rtol=0, atol=1e-18)
    assert np.allclose(d1 - d0, [0, 0, -0.5 * dt * dt * nm.GRAV, 0, 0, 0, 0, 0, 0], rtol=0, atol=1e-18)


@pytest.mark.parametrize("mu", [0.8, 0.3, 0.1])
def test_oracle_solution_satisfies_independent_kkt(pkg, orc, wl, mu):
    cfg = hard_config(wl, 10, mu)
    st, ds, di = wl.make_batch(cfg, 24, gaits=wl.GAITS)
    cc = pkg.make_config(cfg)
    res = orc.solve_batch(cc, st, ds, di, nthreads=4)
    assert (res["status"] == 0).all()
    U = to_step_major(res["forces"], 10, 4)
    nact = 0
    for b in range(24):
        qp = nm.build_qp(cfg, st[b], ds[b], di[b])
        C, lb, ub, tags = nm.constraints(cfg, qp["contact"])
        ll = np.array([res["lam"][b, 0, j, i, r] for (j, i, r) in tags])
        lu = np.array([res["lam"][b, 1, j, i, r] for (j, i, r) in tags])
        assert nm.kkt_residual(cfg, qp, U[b], ll, lu) <= 1e-12
        assert abs(res["kkt"][b] - nm.kkt_residual(cfg, qp, U[b], ll, lu)) <= 1e-12
        # active mask == rows with a positive multiplier or zero slack in the mask
        for t, (j, i, r) in enumerate(tags):
            bit = (int(res["active"][b, j, i]) >> r) & 1
            if ll[t] > 1e-9:
                assert bit == 1
            if bit:
                assert abs((C @ U[b])[t] - lb[t]) <= 1e-9 * (1 + np.abs(U[b]).max())
            nact += bit
        # swing legs are pinned and flagged
        assert ((res["active"][b] == 0x8000) == (qp["contact"].T <= 0)).all()
    assert nact > 20


def test_oracle_agrees_with_textbook_active_set(pkg, orc, wl):
    cfg = hard_config(wl, 6, 0.3)
    st, ds, di = wl.make_batch(cfg, 10, gaits=("trot", "stand"))
    cc = pkg.make_config(cfg)
    res = orc.solve_batch(cc, st, ds, di)
    U = to_step_major(res["forces"], 6, 4)
    checked = 0
    for b in range(10):
        qp = nm.build_qp(cfg, st[b], ds[b], di[b])
        try:
            Ua, _, _, _ = nm.solve_active_set(cfg, qp)
        except RuntimeError:
            continue
        assert np.abs(Ua - U[b]).max() <= 1e-8 * np.abs(Ua).max()
        checked += 1
    assert checked >= 5


def test_known_solution_unconstrained(pkg, orc, wl):
    """knownSolution pattern (testHpipmInterface.cpp:112-152): with no row active the optimum
    solves H U = -g exactly on the free variables."""
    cfg, st, ds, di = wl.fixture_f1_n10()
    cc = pkg.make_config(cfg)
    res = orc.solve_batch(cc, st[None], ds[None], di[None])
    H, g, s = orc.build_qp(cc, st, ds, di)
    U = to_step_major(res["forces"], 10, 4)[0]
    assert (res["active"][0] & 0x3FF == 0).all()
    assert np.abs(H @ U + g).max() <= 1e-12 * (1 + np.abs(g).max())
    assert np.allclose(U, np.linalg.solve(H, -g), rtol=1e-11, atol=1e-11)


def test_cost_from_H_g_equals_stage_cost_rollout(pkg, orc, wl):
    """build-vs-evaluate consistency (testTranscription.cpp:61-65 pattern): 1/2 U'HU + g'U + c0
    equals the reference's stage costs (CentroidalMPC.cpp:208-231) summed along the rollout."""
    cfg = hard_config(wl, 8, 0.5)
    st, ds, di = wl.make_batch(cfg, 3, gaits=("trot", "gallop", "stand"))
    cc = pkg.make_config(cfg)
    rng = np.random.default_rng(1)
    for b in range(3):
        qp = nm.build_qp(cfg, st[b], ds[b], di[b])
        H, g, s = orc.build_qp(cc, st[b], ds[b], di[b])
        U = rng.uniform(-20, 40, H.shape[0]) * (~qp["pinned"])
        quad = 0.5 * U @ H @ U + g @ U + qp["cost0"]
        assert abs(quad - nm.stage_cost(cfg, qp, U)) <= 1e-10 * abs(quad)


def test_value_function_quadratic_while_active_set_fixed(pkg, orc, wl):
    """testValuefunction.cpp:96-105 pattern: optimal forces are affine in x0 while the active
    set does not change (here: no active rows), so second differences vanish."""
    cfg, st, ds, di = wl.fixture_f1_n10()
    cc = pkg.make_config(cfg)
    d = np.zeros(21); d[3] = 0.01; d[8] = 0.02
    F = [orc.solve_batch(cc, (st + k * d)[None], ds[None], di[None])["forces"][0] for k in (-1, 0, 1)]
    assert np.abs(F[0] - 2 * F[1] + F[2]).max() <= 1e-9 * np.abs(F[1]).max()


def test_invalid_table_and_nonfinite(pkg, orc, wl):
    cfg = wl.default_config(10)
    st, ds, di = wl.make_batch(cfg, 4)
    di[1].reshape(4, 43)[:, 0] = 0.0
    st[2, 5] = np.inf
    res = orc.solve_batch(pkg.make_config(cfg), st, ds, di)
    assert res["status"].tolist() == [0, 3, 4, 0]
    assert (res["forces"][1] == 0).all() and (res["forces"][2] == 0).all()


def test_single_leg_and_biped_configs(pkg, orc, wl):
    """num_legs is a constructor argument of the reference (CentroidalMPC.h:26)."""
    for L in (1, 2):
        N = 5
        w = [1, 1, 100, .5, .5, 0, 2, 2, 8] + [0.2] * (3 * L) + [0.3] * (3 * L) + [0.1] * (3 * L)
        cfg = dict(mass=8.0, num_legs=L, horizon=N, dt=0.01, mu=[0.6] * L, weights=w, disc_mode=0)
        st = np.concatenate([[0, 0, 0.15, 0.05, 0, 0, 0, 0, 0], np.tile([0.0, 0.05, 0.0], L)])
        pos = np.array([[0.0005 * k, 0, 0.15] for k in range(N + 1)])
        ds = np.concatenate([pos.ravel(), np.tile([0.05, 0, 0], N + 1), np.zeros(3 * (N + 1))])
        di = wl.pack_des_inputs(np.ones((L, N)), np.tile([0.0, 0.05, 0.0], (L, N + 1, 1)))
        res = orc.solve_batch(pkg.make_config(cfg), st[None], ds[None], di[None])
        assert res["status"][0] == 0 and res["kkt"][0] <= 1e-12
        fz = res["forces"][0].reshape(L, N, 3)[:, :, 2]
        assert np.all(fz > 0) and np.all(fz <= 8 * 9.81 * L + 1e-9)


def test_plant_step_matches_mirror(pkg, orc, wl):
    cfg = wl.default_config(10)
    rng = np.random.default_rng(3)
    x, feet, f = rng.normal(size=9), rng.normal(size=(4, 3)), rng.normal(size=(4, 3)) * 30
    c = np.array([1.0, 0.0, 1.0, 1.0])
    assert np.allclose(orc.plant_step(pkg.make_config(cfg), x, feet, c, f), nm.nonlinear_step(cfg, x, feet, c, f),
                       rtol=1e-14, atol=1e-14)


def test_batch_threads_agree(pkg, orc, wl):
    cfg = wl.default_config(10)
    st, ds, di = wl.make_batch(cfg, 64, gaits=wl.GAITS)
    cc = pkg.make_config(cfg)
    a = orc.solve_batch(cc, st, ds, di, nthreads=1)
    b = orc.solve_batch(cc, st, ds, di, nthreads=5)
    assert np.array_equal(a["forces"], b["forces"]) and np.array_equal(a["status"], b["status"])


@pytest.mark.parametrize("N,disc", [(6, 0), (10, 0), (10, 1), (30, 0)])
def test_stage_wise_riccati_equals_condensed_solve(pkg, wl, N, disc):
    """the un-condensed (stage-wise, Riccati over [x; F_prev]) form has the same
    unconstrained minimiser as the condensed dense H, and its adjoint gradient equals H U + g."""
    import numpy_mirror as nm
    cfg = wl.default_config(N, disc_mode=disc)
    st, ds, di = wl.make_batch(cfg, 5, gaits=wl.GAITS)
    for b in range(5):
        qp = nm.build_qp(cfg, st[b], ds[b], di[b])
        Ud = np.linalg.solve(qp["H"], -qp["g"])
        Ur = nm.riccati_unconstrained(cfg, st[b], ds[b], di[b])
        assert np.abs(Ud - Ur).max() <= 1e-12 * np.abs(Ud).max()
        g0 = nm.stage_gradient(cfg, st[b], ds[b], di[b], np.zeros_like(Ur))
        assert np.abs(g0 - qp["g"]).max() <= 1e-12 * (1 + np.abs(qp["g"]).max())
        rng = np.random.default_rng(b)
        U = rng.normal(size=Ur.shape) * ~qp["pinned"]
        gu = nm.stage_gradient(cfg, st[b], ds[b], di[b], U)
        assert np.abs(gu - (qp["H"] @ U + qp["g"]) * ~qp["pinned"]).max() <= 1e-11 * (1 + np.abs(gu).max())


@pytest.mark.parametrize("N,disc", [(6, 0), (10, 1)])
def test_stage_newton_step_equals_dense_system(wl, N, disc):
    """The stage-wise (Riccati) form of one interior-point step / one polish pass (numpy_mirror.stage_newton_step,
    the mirror of cmpc_ripm.cu) against the dense solves (H + C'SC)^-1 rhs and Z (Z'HZ)^-1 Z' rhs."""
    import numpy_mirror as nm
    from conftest import hard_config
    cfg = hard_config(wl, N, 0.3, disc_mode=disc)
    st, ds, di = wl.make_batch(cfg, 3, gaits=("gallop", "stand", "trot"))
    rng = np.random.default_rng(7)
    L, nf = 4, 12
    for b in range(3):
        qp = nm.build_qp(cfg, st[b], ds[b], di[b])
        C, lb, ub, tags = nm.constraints(cfg, qp["contact"])
        sv = rng.uniform(0.01, 100, len(tags))
        sig = np.zeros((N, L, 5))
        for t, (j, i, r) in enumerate(tags):
            sig[j, i, r] = sv[t]
        free = ~qp["pinned"]
        rhs = rng.normal(size=N * nf) * free
        dref = np.zeros(N * nf)
        dref[free] = np.linalg.solve((qp["H"] + C.T @ (sv[:, None] * C))[np.ix_(free, free)], rhs[free])
        d = nm.stage_newton_step(cfg, st[b], ds[b], di[b], sig, rhs)
        assert np.abs(d - dref).max() <= 1e-12 * np.abs(dref).max()
        Zm, cols = [], []
        for k in range(N):
            Zk = []
            for i in range(L):
                if qp["contact"][i, k] > 0:
                    Q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
                    for c in range(rng.integers(0, 4)):
                        v = np.zeros(nf); v[3 * i:3 * i + 3] = Q[:, c]; Zk.append(v)
                        full = np.zeros(N * nf); full[nf * k:nf * k + nf] = v; cols.append(full)
            Zm.append(np.array(Zk).T.reshape(nf, -1) if Zk else np.zeros((nf, 0)))
        Zf = np.array(cols).T
        dref2 = Zf @ np.linalg.solve(Zf.T @ qp["H"] @ Zf, Zf.T @ rhs)
        d2 = nm.stage_newton_step(cfg, st[b], ds[b], di[b], np.zeros((N, L, 5)), rhs, Zm)
        assert np.abs(d2 - dref2).max() <= 1e-11 * np.abs(dref2).max()


@pytest.mark.parametrize("N,hard,disc", [(10, False, 0), (10, True, 0), (6, True, 1), (30, True, 0), (30, False, 0)])
def test_fast_cpu_port_matches_oracle(pkg, orc, wl, N, hard, disc):
    """The CPU baseline port that bench.py times (oracle/cmpc_cpu_fast.c: compact closed-form build, workspace,
    thread pool) against the explicit oracle: same status, iterations and active set, forces to 1e-9."""
    from conftest import hard_config
    cfg = hard_config(wl, N, 0.3, disc_mode=disc) if hard else wl.default_config(N, disc_mode=disc)
    B = 24 if N == 30 else 96
    st, ds, di = wl.make_batch(cfg, B, gaits=wl.GAITS)
    di[1].reshape(4, 4 * N + 3)[:, 2] = 0.0    # invalid table
    st[2, 4] = np.inf                           # non-finite input
    ccfg = pkg.make_config(cfg)
    ref = orc.solve_batch(ccfg, st, ds, di, nthreads=4)
    for threads in (1, 3):
        out = orc.fast_solve_batch(ccfg, st, ds, di, nthreads=threads)
        assert np.array_equal(out["status"], ref["status"])
        assert out["status"][1] == 3 and out["status"][2] == 4
        assert np.array_equal(out["iters"], ref["iters"])
        assert np.array_equal(out["active"], ref["active"])
        sc = np.abs(ref["forces"]).max(axis=1, keepdims=True) + 1e-300
        assert (np.abs(out["forces"] - ref["forces"]) / sc).max() <= 1e-9
        assert out["kkt"].max() <= 1e-8


def _golden_cfg(v):
    L = int(v[1])
    return dict(mass=float(v[0]), num_legs=L, horizon=int(v[2]), dt=float(v[3]), disc_mode=int(v[4]),
                mu=list(v[5:5 + L]), weights=list(v[5 + L:5 + L + 9 + 9 * L]))


def test_scipy_pin(pkg, orc):
    """The oracle against arithmetic this repo did not write (tests/golden/make_scipy_pin.py): SciPy's SLSQP on the
    swing-eliminated QP of every golden case -- raw SLSQP forces <= 5e-6 relative (SLSQP's own accuracy), its
    active set refined by one LAPACK KKT solve <= 1e-9, identical active set incl. the cases with 28-69 active rows."""
    import os
    from conftest import to_step_major
    here = os.path.join(os.path.dirname(__file__), "golden")
    z, p = np.load(os.path.join(here, "golden_v1.npz")), np.load(os.path.join(here, "scipy_pin_v1.npz"))
    nact = 0
    for name in sorted({k.split("/")[0] for k in z.files}):
        cfg = _golden_cfg(z[name + "/cfg"])
        ref = orc.solve_batch(pkg.make_config(cfg), z[name + "/state"], z[name + "/des_state"], z[name + "/des_inputs"])
        U = to_step_major(ref["forces"], cfg["horizon"], cfg["num_legs"])
        sc = np.abs(U).max(axis=1, keepdims=True)
        assert (np.abs(U - p[name + "/U_slsqp"]) / sc).max() <= 5e-6, name
        assert (np.abs(U - p[name + "/U_refined"]) / sc).max() <= 1e-9, name
        assert np.array_equal(ref["active"] & 0x3FF, p[name + "/active_slsqp"]), name
        nact += sum(bin(int(a)).count("1") for a in p[name + "/active_slsqp"].ravel())
    assert nact > 500   # the pin covers active rows, not only interior optima


def test_scipy_pin_live_small(pkg, orc, wl):
    """Live SciPy run (SLSQP + trust-constr) on the reference driver's fixture: agrees with the oracle."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_scipy_pin as M
    from conftest import to_step_major
    cfg, st, ds, di = wl.fixture_f1()
    H, g, C, ub, tags, free, qp = M.reduced_qp(cfg, st, ds, di)
    u0 = M.feasible_start(cfg, qp, free, C, ub, tags)
    u, _ = M.solve_slsqp(H, g, C, ub, u0)
    u2, _ = M.solve_trust_constr(H, g, C, ub, u0)
    ref = orc.solve_batch(pkg.make_config(cfg), st[None], ds[None], di[None])
    U = to_step_major(ref["forces"], cfg["horizon"], cfg["num_legs"])[0][free]
    assert np.abs(u - U).max() <= 5e-6 * np.abs(U).max()
    assert np.abs(u2 - U).max() <= 5e-6 * np.abs(U).max()


def test_reference_nlp_pin(pkg, orc):
    """Model fidelity: the reference's own non-convex NLP (CentroidalMPC.cpp:102-276) solved by SciPy
    (tests/golden/make_nlp_pin.py) against the convex model.  On the reference driver's fixture the frozen-arm QP's
    forces are within 1 % of the NLP's, the re-linearised fixed point within 0.05 %; with tracking-heavy weights the
    NLP moves the swing footholds to the edge of the step box and the convex model (footholds fixed) is 13-45 % off --
    the stored gaps are re-derived here from the oracle, so they cannot drift silently."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_nlp_pin as M
    p = np.load(os.path.join(os.path.dirname(__file__), "golden", "nlp_pin_v1.npz"))
    for name in sorted({k.split("/")[0] for k in p.files}):
        cfg = _golden_cfg(p[name + "/cfg"])
        st, ds, di = p[name + "/state"], p[name + "/des_state"], p[name + "/des_inputs"]
        seq = M.relinearised_fixed_point(pkg, orc, cfg, st, ds, di)
        fn = p[name + "/forces_nlp"]
        sc = np.abs(fn).max()
        gaps = np.array([np.abs(seq[0] - fn).max() / sc, np.abs(seq[-1] - fn).max() / sc])
        assert np.allclose(gaps, p[name + "/gaps"], rtol=1e-4, atol=1e-9), (name, gaps, p[name + "/gaps"])
        assert np.abs(seq[0] - p[name + "/forces_qp"]).max() <= 1e-9 * sc
        if name.startswith("fixture_f1"):
            assert gaps[0] <= 1e-2 and gaps[1] <= 5e-4, (name, gaps)
        # the NLP solution is feasible for the reference's constraints and not worse than the convex model's point
        assert p[name + "/cost_nlp"][0] <= p[name + "/cost_nlp"][1]
