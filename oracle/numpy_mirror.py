"""NumPy mirror of the condensed centroidal-MPC QP  --  TEST INFRASTRUCTURE ONLY.

This file is an independent, dense, slow restatement of the path in SURVEY.md §8(a).
It exists to cross-check ``oracle/cmpc_oracle.c`` (the C oracle) and to generate /
validate the golden vectors under ``tests/golden``.  Nothing in the product path
(``cheeta-mpc_b200/``) may import it.

PARITY UNPINNED: the reference (``/root/reference/CentroidalMPC.cpp``) returns an empty
vector from ``UpdateMPC`` (:369) and its only driver asserts nothing
(``CentoidMPCTest.cpp:113-115``); CasADi/IPOPT/HSL are absent, so no reference output
exists to pin against.  Every term below cites the reference line it restates.

Layouts (reference ``CentroidalMPC.cpp:278-323``):
  state      [9+3L]      c(3) v(3) Lm(3) p_i(3)...
  des_state  [9(N+1)]    three column-major 3x(N+1) blocks: com_pos | com_vel | ang_mom
  des_inputs [L(4N+3)]   per leg: contact_enable (N) | des_foot_pos 3x(N+1) col-major
Decision vector U in R^{3LN}, step-major: index 3L*j + 3*i + r  (step j, leg i, xyz r).
"""
from __future__ import annotations

import numpy as np

GRAV = 9.81  # CentroidalMPC.cpp:71
FRIC_UB = 5000.0  # CentroidalMPC.cpp:183


def skew(r):
    return np.array([[0.0, -r[2], r[1]], [r[2], 0.0, -r[0]], [-r[1], r[0], 0.0]])


def unpack(cfg, state, des_state, des_inputs):
    """CentroidalMPC.cpp:284-323 memcpy semantics."""
    N, L = cfg["horizon"], cfg["num_legs"]
    state = np.asarray(state, float)
    des_state = np.asarray(des_state, float)
    des_inputs = np.asarray(des_inputs, float)
    x0 = state[:9].copy()
    feet = state[9:9 + 3 * L].reshape(L, 3)
    dc = des_state[0:3 * (N + 1)].reshape(N + 1, 3)          # node k -> xyz
    dv = des_state[3 * (N + 1):6 * (N + 1)].reshape(N + 1, 3)
    dl = des_state[6 * (N + 1):9 * (N + 1)].reshape(N + 1, 3)
    contact = np.zeros((L, N))
    dfoot = np.zeros((L, N + 1, 3))
    for i in range(L):
        blk = des_inputs[i * (4 * N + 3):(i + 1) * (4 * N + 3)]
        contact[i] = blk[:N]
        dfoot[i] = blk[N:].reshape(N + 1, 3)
    return x0, feet, dc, dv, dl, contact, dfoot


def discretize(cfg, contact_j, r_j):
    """A_d, B_j, d for one interval (CentroidalMPC.cpp:85-92, lever arms frozen).

    disc_mode 0: explicit Euler (the reference).  1: zero-order hold, obtained here
    generically from the matrix exponential of the augmented continuous system
    (series terminates: the augmented matrix is nilpotent of index 3).
    """
    m, dt, L = cfg["mass"], cfg["dt"], cfg["num_legs"]
    Ac = np.zeros((9, 9))
    Ac[0:3, 3:6] = np.eye(3)
    Bc = np.zeros((9, 3 * L))
    for i in range(L):
        c = contact_j[i] if contact_j[i] > 0 else 0.0
        Bc[3:6, 3 * i:3 * i + 3] = c / m * np.eye(3)
        Bc[6:9, 3 * i:3 * i + 3] = c * skew(r_j[i])
    dc_ = np.zeros(9)
    dc_[5] = -GRAV
    if cfg.get("disc_mode", 0) == 0:
        return np.eye(9) + dt * Ac, dt * Bc, dt * dc_
    nz = 9 + 3 * L + 1
    M = np.zeros((nz, nz))
    M[:9, :9] = Ac * dt
    M[:9, 9:9 + 3 * L] = Bc * dt
    M[:9, -1] = dc_ * dt
    E = np.eye(nz) + M + M @ M / 2.0 + M @ M @ M / 6.0
    return E[:9, :9], E[:9, 9:9 + 3 * L], E[:9, -1]


def build_qp(cfg, state, des_state, des_inputs):
    """Dense H (p x p), g (p), plus pieces.  Returns dict.  SURVEY §8 a2-a7."""
    N, L, m = cfg["horizon"], cfg["num_legs"], cfg["mass"]
    w = np.asarray(cfg["weights"], float)
    nu, p, q = 3 * L, 3 * L * N, 9 * N
    x0, feet, dc, dv, dl, contact, dfoot = unpack(cfg, state, des_state, des_inputs)
    colsum = contact.sum(axis=0)
    invalid = bool(np.any(colsum <= 0))  # CentroidalMPC.cpp:328-330
    stance = contact > 0

    Aqp = np.zeros((q, 9))
    Bqp = np.zeros((q, p))
    dqp = np.zeros(q)
    Ad = None
    Bs, ds = [], []
    for j in range(N):
        r_j = dfoot[:, j, :] - dc[j][None, :]
        Ad, Bj, dj = discretize(cfg, contact[:, j], r_j)
        Bs.append(Bj)
        ds.append(dj)
    Apow = [np.eye(9)]
    for k in range(N):
        Apow.append(Ad @ Apow[-1])
    for k in range(N):
        Aqp[9 * k:9 * k + 9] = Apow[k + 1]
        acc = np.zeros(9)
        for j in range(k + 1):
            Bqp[9 * k:9 * k + 9, nu * j:nu * j + nu] = Apow[k - j] @ Bs[j]
            acc += Apow[k - j] @ ds[j]
        dqp[9 * k:9 * k + 9] = acc

    # cost weights, CentroidalMPC.cpp:203-216 (omega_k inside the square)
    Lw = np.zeros(q)
    Xref = np.zeros(q)
    for k in range(N):
        node = k + 1
        om = (w[2] / 2.0) * np.exp(-float(node)) + w[2] / 2.0
        Lw[9 * k:9 * k + 9] = [w[0], w[1], om * om, w[3], w[4], w[5], w[6], w[7], w[8]]
        Xref[9 * k:9 * k + 9] = np.concatenate([dc[node], dv[node], dl[node]])
    # force tracking + rate, CentroidalMPC.cpp:223-231
    wf = np.zeros(p)
    Uref = np.zeros(p)
    D = np.zeros((max(N - 1, 0) * nu, p))
    wr = np.zeros(max(N - 1, 0) * nu)
    for j in range(N):
        for i in range(L):
            for r in range(3):
                wf[nu * j + 3 * i + r] = w[9 + 3 * L + 3 * i + r]
            if stance[i, j] and not invalid:
                Uref[nu * j + 3 * i + 2] = m * GRAV / colsum[j]  # :331-333
    for j in range(N - 1):
        for a in range(nu):
            D[nu * j + a, nu * j + a] = -1.0
            D[nu * j + a, nu * (j + 1) + a] = 1.0
            i, r = divmod(a, 3)
            wr[nu * j + a] = w[9 + 6 * L + 3 * i + r]
    K = np.diag(wf) + D.T @ (wr[:, None] * D)
    H = 2.0 * (Bqp.T @ (Lw[:, None] * Bqp) + K)
    xfree = Aqp @ x0 + dqp
    g = 2.0 * Bqp.T @ (Lw * (xfree - Xref)) - 2.0 * wf * Uref
    cost0 = float((xfree - Xref) @ (Lw * (xfree - Xref)) + Uref @ (wf * Uref))

    pinned = np.ones(p, bool)
    for j in range(N):
        for i in range(L):
            if stance[i, j]:
                pinned[nu * j + 3 * i:nu * j + 3 * i + 3] = False
    Hm, gm = H.copy(), g.copy()
    Hm[pinned, :] = 0.0
    Hm[:, pinned] = 0.0
    Hm[pinned, pinned] = 1.0
    gm[pinned] = 0.0
    return dict(H=Hm, g=gm, H_raw=H, g_raw=g, pinned=pinned, invalid=invalid, Aqp=Aqp,
                Bqp=Bqp, dqp=dqp, Lw=Lw, Xref=Xref, K=K, Uref=Uref, wf=wf, x0=x0,
                contact=contact, stance=stance, cost0=cost0)


def constraints(cfg, contact):
    """Dense C (5 rows per stance leg-step), lb, ub over the FULL U layout.
    CentroidalMPC.cpp:179-200.  Returns C, lb, ub, rows -> (j, i, r)."""
    N, L, m = cfg["horizon"], cfg["num_legs"], cfg["mass"]
    nu = 3 * L
    rowsC, lb, ub, tags = [], [], [], []
    for j in range(N):
        for i in range(L):
            c = contact[i, j]
            if c <= 0:
                continue
            mu = cfg["mu"][i]
            F = np.array([[-1, 0, mu], [1, 0, mu], [0, -1, mu], [0, 1, mu], [0, 0, 1.0]])
            ubv = c * np.array([FRIC_UB] * 4 + [m * GRAV * L])
            for r in range(5):
                row = np.zeros(nu * N)
                row[nu * j + 3 * i:nu * j + 3 * i + 3] = F[r]
                rowsC.append(row)
                lb.append(0.0)
                ub.append(ubv[r])
                tags.append((j, i, r))
    return np.array(rowsC), np.array(lb), np.array(ub), tags


def solve_active_set(cfg, qp, max_iter=2000, tol=1e-11):
    """Independent solver: primal active-set on the free variables (textbook,
    Nocedal & Wright alg. 16.3) started from the interior point f=(0,0,fz_ref).
    Slow; used only to pin the IPM oracle.  Returns U (full layout), lam_l, lam_u, tags."""
    p = qp["H"].shape[0]
    free = np.where(~qp["pinned"])[0]
    H = qp["H"][np.ix_(free, free)]
    g = qp["g"][free]
    C, lb, ub, tags = constraints(cfg, qp["contact"])
    C = C[:, free]
    mrows = C.shape[0]
    # feasible start
    U0 = np.zeros(p)
    N, L = cfg["horizon"], cfg["num_legs"]
    for j in range(N):
        for i in range(L):
            if qp["stance"][i, j]:
                c = qp["contact"][i, j]
                fzmax = c * cfg["mass"] * GRAV * L
                U0[3 * L * j + 3 * i + 2] = min(0.5 * fzmax, 0.5 * c * FRIC_UB / cfg["mu"][i])
    u = U0[free]
    # working set W: list of (row, side) side=+1 lower (C u >= lb), -1 upper
    W = []
    for it in range(max_iter):
        if W:
            A = np.array([s * C[r] for r, s in W])
            KKT = np.block([[H, -A.T], [A, np.zeros((len(W), len(W)))]])
            rhs = np.concatenate([-(H @ u + g), np.zeros(len(W))])
            sol = np.linalg.lstsq(KKT, rhs, rcond=None)[0]
            d, lam = sol[:len(u)], sol[len(u):]
        else:
            d = np.linalg.solve(H, -(H @ u + g))
            lam = np.zeros(0)
        if np.linalg.norm(d, np.inf) < tol * (1 + np.linalg.norm(u, np.inf)):
            if len(W) == 0 or lam.min() >= -1e-10:
                break
            W.pop(int(np.argmin(lam)))
            continue
        alpha, blk = 1.0, None
        Cd, Cu = C @ d, C @ u
        inW = set(W)
        for r in range(mrows):
            if (r, 1) not in inW and Cd[r] < -1e-14:
                a = (lb[r] - Cu[r]) / Cd[r]
                if a < alpha:
                    alpha, blk = max(a, 0.0), (r, 1)
            if (r, -1) not in inW and Cd[r] > 1e-14:
                a = (ub[r] - Cu[r]) / Cd[r]
                if a < alpha:
                    alpha, blk = max(a, 0.0), (r, -1)
        u = u + alpha * d
        if blk is not None:
            # keep the working set linearly independent
            A = np.array([s * C[r] for r, s in W] + [blk[1] * C[blk[0]]])
            if np.linalg.matrix_rank(A, tol=1e-10) == len(W) + 1:
                W.append(blk)
    else:
        raise RuntimeError("active-set did not converge")
    U = np.zeros(p)
    U[free] = u
    lam_l = np.zeros(mrows)
    lam_u = np.zeros(mrows)
    for (r, s), l in zip(W, lam):
        if s > 0:
            lam_l[r] = l
        else:
            lam_u[r] = l
    return U, lam_l, lam_u, tags


def kkt_residual(cfg, qp, U, lam_l, lam_u):
    """Scaled KKT residual of (U, lam) on the masked QP; definition shared with the C
    oracle and the CUDA path."""
    C, lb, ub, _ = constraints(cfg, qp["contact"])
    H, g = qp["H"], qp["g"]
    if C.shape[0] == 0:
        return float(np.abs(H @ U + g).max() / (1 + np.abs(g).max()))
    r = H @ U + g - C.T @ lam_l + C.T @ lam_u
    sl, su = C @ U - lb, ub - C @ U
    gs, us = 1 + np.abs(g).max(), 1 + np.abs(U).max()
    stat = np.abs(r).max() / gs
    prim = max(0.0, (-sl).max(), (-su).max()) / us
    dual = max(0.0, (-lam_l).max(), (-lam_u).max()) / gs
    comp = max(np.abs(lam_l * sl).max(), np.abs(lam_u * su).max()) / (gs * us)
    return float(max(stat, prim, dual, comp))


def nonlinear_step(cfg, x, feet, contact_j, forces_j):
    """Reference plant, verbatim Euler (CentroidalMPC.cpp:85-92); feet fixed."""
    m, dt, L = cfg["mass"], cfg["dt"], cfg["num_legs"]
    c, v, lm = x[0:3], x[3:6], x[6:9]
    acc = np.array([0.0, 0.0, -GRAV])
    ldot = np.zeros(3)
    for i in range(L):
        acc = acc + contact_j[i] / m * forces_j[i]
        ldot = ldot + contact_j[i] * np.cross(feet[i] - c, forces_j[i])
    return np.concatenate([c + v * dt, v + acc * dt, lm + ldot * dt])


def stage_cost(cfg, qp, U):
    """Cost by rolling the *linear* condensed dynamics forward and summing the
    reference's stage costs (CentroidalMPC.cpp:208-231) -- build-vs-evaluate test."""
    X = qp["Aqp"] @ qp["x0"] + qp["Bqp"] @ U + qp["dqp"]
    e = X - qp["Xref"]
    N, L = cfg["horizon"], cfg["num_legs"]
    nu = 3 * L
    w = np.asarray(cfg["weights"], float)
    J = float(e @ (qp["Lw"] * e))
    J += float((U - qp["Uref"]) @ (qp["wf"] * (U - qp["Uref"])))
    for j in range(N - 1):
        for a in range(nu):
            i, r = divmod(a, 3)
            J += w[9 + 6 * L + 3 * i + r] * (U[nu * (j + 1) + a] - U[nu * j + a]) ** 2
    return J


# ---------------------------------------------------------------------------- gait
# ModeNumber bits {LF=8, RF=4, LH=2, RH=1}: ocs2_legged_robot/include/ocs2_legged_robot/gait/
# MotionPhaseDefinition.h:47-64,129-132.  Templates: ocs2_legged_robot/config/command/gait.info.
MODE = dict(FLY=0, RH=1, LH=2, LH_RH=3, RF=4, RF_RH=5, RF_LH=6, RF_LH_RH=7, LF=8, LF_RH=9, LF_LH=10,
            LF_LH_RH=11, LF_RF=12, LF_RF_RH=13, LF_RF_LH=14, STANCE=15)
GAIT_INFO = {  # name -> (modeSequence, switchingTimes), values of gait.info:17-180
    "stance": (["STANCE"], [0.0, 0.5]),
    "trot": (["LF_RH", "RF_LH"], [0.0, 0.35, 0.70]),
    "standing_trot": (["LF_RH", "STANCE", "RF_LH", "STANCE"], [0.00, 0.30, 0.35, 0.65, 0.70]),
    "flying_trot": (["LF_RH", "FLY", "RF_LH", "FLY"], [0.00, 0.27, 0.30, 0.57, 0.60]),
    "pace": (["LF_LH", "FLY", "RF_RH", "FLY"], [0.0, 0.28, 0.30, 0.58, 0.60]),
    "standing_pace": (["LF_LH", "STANCE", "RF_RH", "STANCE"], [0.0, 0.30, 0.35, 0.65, 0.70]),
    "dynamic_walk": (["LF_RF_RH", "RF_RH", "RF_LH_RH", "LF_RF_LH", "LF_LH", "LF_LH_RH"], [0.0, 0.2, 0.3, 0.5, 0.7, 0.8, 1.0]),
    "static_walk": (["LF_RF_RH", "RF_LH_RH", "LF_RF_LH", "LF_LH_RH"], [0.0, 0.3, 0.6, 0.9, 1.2]),
    "amble": (["RF_LH", "LF_LH", "LF_RH", "RF_RH"], [0.0, 0.15, 0.40, 0.55, 0.80]),
}


def gait_contact_table(mode_sequence, switching_times, t0, dt, N, L=4):
    """Contact table [L, N] in the reference driver's leg order {lf, rf, rh, lh}
    (CentoidMPCTest.cpp:43-46).  Restates toGait (ModeSequenceTemplate.cpp:74-87), wrapPhase and
    getModeFromPhase (Gait.cpp:63-88, std::upper_bound) and modeNumber2StanceLeg."""
    st = np.asarray(switching_times, float)
    start, duration = st[0], st[-1] - st[0]
    event_phases = (st[1:-1] - start) / duration
    modes = [MODE[m] if isinstance(m, str) else int(m) for m in mode_sequence]
    bit = [8, 4, 1, 2]  # {lf, rf, rh, lh} <- {LF, RF, RH, LH}
    out = np.zeros((L, N))
    for j in range(N):
        phase = np.fmod((t0 + j * dt) / duration, 1.0)
        if phase < 0:
            phase += 1.0
        k = int(np.searchsorted(event_phases, phase, side="right"))
        for i in range(L):
            out[i, j] = 1.0 if modes[k] & bit[i] else 0.0
    return out


def gait_switch_contact_table(gait_from, gait_to, t_tile, t_switch, stance_time, t0, dt, N, L=4):
    """Contact table [L, N] of a mode schedule in which template gait_from = (mode_sequence, switching_times), tiled from
    t_tile, is replaced at t_switch by gait_to with an intermediate stance phase.  The schedule is built EXPLICITLY, the
    way the reference stack does: GaitSchedule::tileModeSequenceTemplate (GaitSchedule.cpp:107-137) and
    GaitSchedule::insertModeSequenceTemplate (:47-72); mode at time t = modeSequence[lower_bound(eventTimes, t)]
    (ocs2_core ModeSchedule::modeAtTime).  t_switch None / NaN: no switch."""
    STANCE = 15

    def tile(events, modes, tmpl, start, final):
        seq, times = tmpl
        seq = [MODE[x] if isinstance(x, str) else int(x) for x in seq]
        if not seq:
            return
        if events and start <= events[-1]:
            raise RuntimeError("The initial time for template-tiling is not greater than the last event time.")
        events.append(start)
        while events[-1] < final:
            for i in range(len(seq)):
                modes.append(seq[i])
                events.append(events[-1] + (times[i + 1] - times[i]))
        modes.append(STANCE)

    t_end = t0 + N * dt
    sw = t_switch is not None and np.isfinite(t_switch)
    events, modes = [], [STANCE]
    tile(events, modes, gait_from, float(t_tile), float(t_switch) if sw else t_end)
    if sw:
        index = int(np.searchsorted(events, t_switch, side="left"))
        if index < len(events):
            del events[index:]
            del modes[index + 1:]
        T = stance_time
        if modes and modes[-1] == STANCE:
            T = 0.0
        if T > 0.0:
            events.append(float(t_switch)); modes.append(STANCE)
        tile(events, modes, gait_to, t_switch + T, t_end)
    bit = [8, 4, 1, 2]  # {lf, rf, rh, lh} <- {LF, RF, RH, LH}
    out = np.zeros((L, N))
    for j in range(N):
        k = int(np.searchsorted(events, t0 + j * dt, side="left"))
        for i in range(L):
            out[i, j] = 1.0 if modes[k] & bit[i] else 0.0
    return out


# ---------------------------------------------------------------------------- foot plan (a3 / a8)
STEP_LB = np.array([-0.2, -0.2, -0.1])  # CentroidalMPC.cpp:30
STEP_UB = np.array([0.2, 0.2, 0.1])     # CentroidalMPC.cpp:31


def foot_plan(cfg, state, des_inputs):
    """Optimal foot positions [L, N+1, 3] of the decoupled foot sub-problem of the reference NLP:
    foot_pos[:,0] = current foot (CentroidalMPC.cpp:166); foot_pos[:,k+1] = foot_pos[:,k] when
    1 - contact_k == 0 (:94) else free; cost w_p sum_k (p_k - pd_k)^2 (:219-221); box on
    p_k - pd_k for k = 1..N (:198).  Nodes joined by locked intervals share one value."""
    N, L = cfg["horizon"], cfg["num_legs"]
    state = np.asarray(state, float)
    D = np.asarray(des_inputs, float).reshape(L, 4 * N + 3)
    out = np.zeros((L, N + 1, 3))
    for i in range(L):
        c, pd = D[i, :N], D[i, N:].reshape(N + 1, 3)
        a = 0
        while a <= N:
            e = a
            while e < N and 1.0 - c[e] == 0.0:
                e += 1
            if a == 0:
                v = state[9 + 3 * i:12 + 3 * i]
            else:
                grp = pd[a:e + 1]
                v = np.minimum(np.maximum(grp.sum(axis=0) / (e - a + 1), (grp + STEP_LB).max(axis=0)), (grp + STEP_UB).min(axis=0))
            out[i, a:e + 1] = v
            a = e + 1
    return out


def foot_cost_and_feasible(cfg, state, des_inputs, p):
    """Cost and feasibility of a candidate foot trajectory p [L, N+1, 3] under the reference's
    constraints -- used to check foot_plan against perturbations."""
    N, L = cfg["horizon"], cfg["num_legs"]
    w = np.asarray(cfg["weights"], float)
    D = np.asarray(des_inputs, float).reshape(L, 4 * N + 3)
    cost, ok = 0.0, True
    for i in range(L):
        c, pd = D[i, :N], D[i, N:].reshape(N + 1, 3)
        cost += float((w[9 + 3 * i:12 + 3 * i] * (p[i] - pd) ** 2).sum())
        ok &= bool(np.allclose(p[i, 0], np.asarray(state, float)[9 + 3 * i:12 + 3 * i], atol=1e-15))
        for k in range(N):
            if 1.0 - c[k] == 0.0:
                ok &= bool(np.array_equal(p[i, k + 1], p[i, k]))
        dlt = p[i, 1:] - pd[1:]
        ok &= bool((dlt >= STEP_LB - 1e-12).all() and (dlt <= STEP_UB + 1e-12).all())
    return cost, ok


# ------------------------------------------------------------------ stage-wise (Riccati) form
def stage_data(cfg, state, des_state, des_inputs):
    """The un-condensed optimal-control form of the same QP (the stage-wise A, B, b, Q, R, q, r
    interface of the reference's HPIPM adapter, ocs2_sqp/hpipm_catkin/src/HpipmInterface.cpp:166-301):
      x_{j+1} = A x_j + Bf_j F_j + d,  F_j in R^{3L} with swing-leg entries pinned to zero,
      cost = sum_{k=1..N} (x_k - xr_k)' Q_k (x_k - xr_k) + sum_j (F_j - Fr_j)' Wf (F_j - Fr_j)
             + sum_{j<=N-2} (F_{j+1} - F_j)' Wr (F_{j+1} - F_j)          (CentroidalMPC.cpp:203-232)."""
    N, L, m = cfg["horizon"], cfg["num_legs"], cfg["mass"]
    w = np.asarray(cfg["weights"], float)
    x0, feet, dc, dv, dl, contact, dfoot = unpack(cfg, state, des_state, des_inputs)
    colsum = contact.sum(axis=0)
    A = None
    Bf, Q, xr, Fr, free = [], [], [], [], []
    for j in range(N):
        A, Bj, d = discretize(cfg, contact[:, j], dfoot[:, j, :] - dc[j][None, :])
        Bf.append(Bj)
        node = j + 1
        om = (w[2] / 2.0) * np.exp(-float(node)) + w[2] / 2.0
        Q.append(np.diag([w[0], w[1], om * om, w[3], w[4], w[5], w[6], w[7], w[8]]))
        xr.append(np.concatenate([dc[node], dv[node], dl[node]]))
        fr = np.zeros(3 * L)
        fj = []
        for i in range(L):
            if contact[i, j] > 0:
                fr[3 * i + 2] = m * GRAV / colsum[j]
                fj += [3 * i, 3 * i + 1, 3 * i + 2]
        Fr.append(fr)
        free.append(np.array(fj, int))
    Wf = np.diag(w[9 + 3 * L:9 + 6 * L])
    Wr = np.diag(w[9 + 6 * L:9 + 9 * L])
    return dict(A=A, d=d, Bf=Bf, Q=Q, xr=xr, Fr=Fr, free=free, Wf=Wf, Wr=Wr, x0=x0, N=N, L=L)


def riccati_unconstrained(cfg, state, des_state, des_inputs):
    """Unconstrained minimiser of the QP by a backward Riccati sweep over the augmented state
    z_k = [x_k; F_{k-1}] (the force-rate term couples consecutive inputs) and a forward roll-out:
    O(N (9 + 3L)^3) instead of O((3LN)^3).  Returns U (3LN, step-major, zeros on pinned entries)."""
    S = stage_data(cfg, state, des_state, des_inputs)
    N, L = S["N"], S["L"]
    nx, nf = 9, 3 * L
    nz = nx + nf
    A, d, Wf, Wr = S["A"], S["d"], S["Wf"], S["Wr"]
    Abar = np.zeros((nz, nz)); Abar[:nx, :nx] = A
    dbar = np.concatenate([d, np.zeros(nf)])
    P = np.zeros((nz, nz)); p = np.zeros(nz)
    P[:nx, :nx] = S["Q"][N - 1]; p[:nx] = -S["Q"][N - 1] @ S["xr"][N - 1]      # V_N
    gains = [None] * N
    for k in range(N - 1, -1, -1):
        fr = S["free"][k]
        Sk = np.zeros((nf, len(fr))); Sk[fr, np.arange(len(fr))] = 1.0
        Bbar = np.vstack([S["Bf"][k] @ Sk, Sk])
        rate = 1.0 if k >= 1 else 0.0
        G = Sk.T @ (Wf + rate * Wr) @ Sk + Bbar.T @ P @ Bbar
        M = Bbar.T @ P @ Abar
        M[:, nx:] -= rate * (Sk.T @ Wr)
        m0 = Bbar.T @ (P @ dbar + p) - Sk.T @ Wf @ S["Fr"][k]
        Kk = np.linalg.solve(G, M); kff = np.linalg.solve(G, m0)
        gains[k] = (Kk, kff, Sk)
        if k >= 1:
            Pn = Abar.T @ P @ Abar - M.T @ Kk
            Pn[:nx, :nx] += S["Q"][k - 1]
            Pn[nx:, nx:] += Wr
            pn = Abar.T @ (P @ dbar + p) - M.T @ kff
            pn[:nx] -= S["Q"][k - 1] @ S["xr"][k - 1]
            P, p = 0.5 * (Pn + Pn.T), pn
    x = S["x0"].copy(); Fp = np.zeros(nf)
    U = np.zeros(nf * N)
    for k in range(N):
        Kk, kff, Sk = gains[k]
        u = -Kk @ np.concatenate([x, Fp]) - kff
        F = Sk @ u
        U[nf * k:nf * (k + 1)] = F
        x = A @ x + S["Bf"][k] @ F + d
        Fp = F
    return U


def stage_gradient(cfg, state, des_state, des_inputs, U):
    """(H U + g) on the free entries by one roll-out and one adjoint (costate) sweep -- independent
    of both the condensed H and the Riccati recursion."""
    S = stage_data(cfg, state, des_state, des_inputs)
    N, L = S["N"], S["L"]
    nf = 3 * L
    F = U.reshape(N, nf)
    xs = [S["x0"]]
    for k in range(N):
        xs.append(S["A"] @ xs[-1] + S["Bf"][k] @ F[k] + S["d"])
    lam = np.zeros(9)
    grad = np.zeros((N, nf))
    for k in range(N - 1, -1, -1):
        lam = 2.0 * S["Q"][k] @ (xs[k + 1] - S["xr"][k]) + S["A"].T @ lam     # costate of x_{k+1}
        gk = 2.0 * S["Wf"] @ (F[k] - S["Fr"][k]) + S["Bf"][k].T @ lam
        if k >= 1:
            gk += 2.0 * S["Wr"] @ (F[k] - F[k - 1])
        if k <= N - 2:
            gk -= 2.0 * S["Wr"] @ (F[k + 1] - F[k])
        mask = np.zeros(nf); mask[S["free"][k]] = 1.0
        grad[k] = gk * mask
    return grad.reshape(-1)


# ------------------------------------------------------------------ stage-wise interior-point step
def stage_newton_step(cfg, state, des_state, des_inputs, sigma, rhs, Zmaps=None):
    """The linear system of one interior-point iteration (or of one active-set polish pass) in the
    stage-wise form the kernel cmpc_ripm.cu uses, restated densely per stage:

        minimise  1/2 d'(H + C' diag(sigma) C) d - rhs' d     over  d_k = Z_k t_k,

    H the condensed Hessian, sigma [N, L, 5] the barrier weights z_l/s_l + z_u/s_u of the friction
    rows (CentroidalMPC.cpp:186-190), Z_k an nf x m_k basis per stage (default: the unit vectors of the
    stance-leg components; the polish passes the per-leg null-space bases of its working set).
    Homogeneous LQR over z_k = [xi_k; d_{k-1}] (deviation state, xi_0 = 0):  backward sweep
    G = Bbar' P Bbar + Z'(Wf + rate Wr + 1/2 C' S C) Z,  M = Bbar' P Abar - rate Z' Wr [0 I],
    Y = L^-1 M, y0 = L^-1 (Bbar' p - Z' rhs_k / 2),  P <- blkdiag(Q_k, Wr) + Abar' P Abar - Y'Y,
    p <- Abar' p - Y' y0;  forward  t_k = -L^-T (Y z_k + y0).  Returns d [N * nf]."""
    S = stage_data(cfg, state, des_state, des_inputs)
    N, L = S["N"], S["L"]
    nx, nf = 9, 3 * L
    nz = nx + nf
    A, Wf, Wr = S["A"], S["Wf"], S["Wr"]
    mu = np.asarray(cfg["mu"], float)
    sigma = np.asarray(sigma, float).reshape(N, L, 5)
    rhs = np.asarray(rhs, float).reshape(N, nf)
    Abar = np.zeros((nz, nz)); Abar[:nx, :nx] = A
    P = np.zeros((nz, nz)); p = np.zeros(nz)
    P[:nx, :nx] = S["Q"][N - 1]
    fac = [None] * N
    for k in range(N - 1, -1, -1):
        if Zmaps is None:
            fr = S["free"][k]
            Z = np.zeros((nf, len(fr))); Z[fr, np.arange(len(fr))] = 1.0
        else:
            Z = np.asarray(Zmaps[k], float).reshape(nf, -1)
        rate = 1.0 if k >= 1 else 0.0
        Rk = Wf + rate * Wr
        for i in range(L):
            Fi = np.array([[-1, 0, mu[i]], [1, 0, mu[i]], [0, -1, mu[i]], [0, 1, mu[i]], [0, 0, 1.0]])
            Rk[3 * i:3 * i + 3, 3 * i:3 * i + 3] += 0.5 * Fi.T @ (sigma[k, i][:, None] * Fi)
        Bbar = np.vstack([S["Bf"][k] @ Z, Z])
        G = Bbar.T @ P @ Bbar + Z.T @ Rk @ Z
        M = Bbar.T @ P @ Abar
        M[:, nx:] -= rate * (Z.T @ Wr)
        m0 = Bbar.T @ p - 0.5 * Z.T @ rhs[k]
        if Z.shape[1]:
            Lc = np.linalg.cholesky(G)
            Y = np.linalg.solve(Lc, M); y0 = np.linalg.solve(Lc, m0)
        else:
            Lc = np.zeros((0, 0)); Y = np.zeros((0, nz)); y0 = np.zeros(0)
        fac[k] = (Lc, Y, y0, Z)
        if k >= 1:
            Pn = Abar.T @ P @ Abar - Y.T @ Y
            Pn[:nx, :nx] += S["Q"][k - 1]
            Pn[nx:, nx:] += Wr
            p = Abar.T @ p - Y.T @ y0
            P = 0.5 * (Pn + Pn.T)
    z = np.zeros(nz)
    d = np.zeros((N, nf))
    for k in range(N):
        Lc, Y, y0, Z = fac[k]
        t = -np.linalg.solve(Lc.T, Y @ z + y0) if Z.shape[1] else np.zeros(0)
        d[k] = Z @ t
        z = np.concatenate([A @ z[:nx] + S["Bf"][k] @ d[k], d[k]])
    return d.reshape(-1)
