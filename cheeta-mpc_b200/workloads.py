"""Synthetic workloads for the batched centroidal MPC (BASELINE.json configs).

Host-side generation with a counter-based RNG (NumPy Philox, key = seed, one
independent stream per instance id) so every rank / every backend sees the same
instance i regardless of how the batch is sharded.

Layouts follow the reference's ``UpdateMPC`` argument packing
(``/root/reference/CentroidalMPC.cpp:284-317``); fixture values follow
``/root/reference/CentoidMPCTest.cpp:12-107``.
"""
from __future__ import annotations

import numpy as np

SEED = 0xC0FFEE

# CentoidMPCTest.cpp:19-33 -- 45 weights, indexed by the code (CentroidalMPC.cpp:208-231)
F1_WEIGHTS = np.array(
    [1, 1, 100, 0.5, 0.5, 0, 2, 2, 8]
    + [0.2, 0.2, 0.2, 0.3, 0.3, 0.3, 0.1, 0.1, 0.1] * 4, dtype=np.float64)
F1_MU = np.array([0.8, 0.8, 0.8, 0.8])
F1_MASS = 8.0
F1_DT = 0.01
# CentoidMPCTest.cpp:40-46 (leg order lf, rf, rh, lh)
F1_STATE = np.array([0, 0, 0.15, 0.1, 0, 0, 0, 0, 0.1,
                     0.35, 0.052, 0, 0.35, -0.054, 0, -0.37, -0.053, 0, -0.36, 0.054, 0],
                    dtype=np.float64)

GAITS = ("stand", "trot", "pace", "bound", "gallop")


def default_config(horizon=10, num_legs=4, mass=F1_MASS, dt=F1_DT, mu=None, weights=None,
                   disc_mode=0):
    return dict(mass=float(mass), num_legs=int(num_legs), horizon=int(horizon), dt=float(dt),
                mu=list(F1_MU[:num_legs] if mu is None else mu),
                weights=list(F1_WEIGHTS if weights is None else weights),
                disc_mode=int(disc_mode))


def sizes(cfg):
    N, L = cfg["horizon"], cfg["num_legs"]
    return dict(state=9 + 3 * L, des_state=9 * (N + 1), des_inputs=L * (4 * N + 3),
                forces=3 * L * N)


def fixture_f1():
    """F1: the reference driver's inputs exactly as its memcpy semantics produce them
    (N=6; des_state has 63 slots of which the comma initialiser fills 54, the last 9
    stay zero -- CentoidMPCTest.cpp:37,48-65; blocks are then read at stride 3(N+1)=21,
    CentroidalMPC.cpp:297-299)."""
    N = 6
    cfg = default_config(horizon=N)
    des_state = np.zeros(9 * (N + 1))
    given = [0.31, 0, 0.16, 0.32, 0, 0.168, 0.33, 0, 0.172, 0.33, 0, 0.18, 0.34, 0, 0.19,
             0.348, 0, 0.2,
             0.1, 0, 0, 0.09, 0, 0, 0.08, 0, 0, 0.06, 0, 0, 0.04, 0, 0, 0, 0, 0,
             0, 0, 0.12, 0, 0, 0.14, 0, 0, 0.16, 0, 0, 0.18, 0, 0, 0.2, 0, 0, 0.22]
    des_state[:len(given)] = given
    table = np.array([[1, 0, 1, 0]] * 3 + [[0, 1, 0, 1]] * 3, dtype=np.float64)  # [j, leg]
    feet = _f1_des_feet()
    des_inputs = pack_des_inputs(table.T, feet)
    return cfg, F1_STATE.copy(), des_state, des_inputs


def _f1_des_feet():
    # CentoidMPCTest.cpp:76-107, [leg][node][xyz]
    f = np.zeros((4, 7, 3))
    f[0] = [[0.35, 0.052, 0]] * 4 + [[0.38, 0.052, 0], [0.39, 0.052, 0], [0.42, 0.052, 0]]
    f[1] = [[0.35, -0.054, 0], [0.37, -0.052, 0], [0.39, -0.052, 0]] + [[0.43, -0.052, 0]] * 4
    f[2] = [[-0.37, -0.052, 0]] * 3 + [[-0.36, -0.052, 0], [-0.34, -0.052, 0],
                                       [-0.30, -0.052, 0], [-0.28, -0.052, 0]]
    f[3] = [[-0.36, 0.053, 0], [-0.34, 0.053, 0], [-0.32, 0.053, 0], [-0.31, 0.053, 0]] \
        + [[-0.31, 0.052, 0]] * 3
    return f


def fixture_f1_intended():
    """F1': the same problem with des_state blocks sized as the driver's comments
    intend (one xyz triplet per node k=0..6 for each of pos/vel/ang-mom; node 0 = the
    current state)."""
    N = 6
    cfg = default_config(horizon=N)
    pos = [[0, 0, 0.15], [0.31, 0, 0.16], [0.32, 0, 0.168], [0.33, 0, 0.172], [0.33, 0, 0.18],
           [0.34, 0, 0.19], [0.348, 0, 0.2]]
    vel = [[0.1, 0, 0], [0.1, 0, 0], [0.09, 0, 0], [0.08, 0, 0], [0.06, 0, 0], [0.04, 0, 0],
           [0, 0, 0]]
    am = [[0, 0, 0.1], [0, 0, 0.12], [0, 0, 0.14], [0, 0, 0.16], [0, 0, 0.18], [0, 0, 0.2],
          [0, 0, 0.22]]
    des_state = np.concatenate([np.ravel(pos), np.ravel(vel), np.ravel(am)]).astype(np.float64)
    table = np.array([[1, 0, 1, 0]] * 3 + [[0, 1, 0, 1]] * 3, dtype=np.float64)
    des_inputs = pack_des_inputs(table.T, _f1_des_feet())
    return cfg, F1_STATE.copy(), des_state, des_inputs


def fixture_f1_n10():
    """Config 1's synthetic N=10 extension: 5 steps legs(0,2) then 5 steps legs(1,3),
    constant-velocity reference, feet held."""
    N = 10
    cfg = default_config(horizon=N)
    vd = np.array([0.1, 0.0, 0.0])
    c0 = F1_STATE[0:3]
    pos = np.array([[c0[0] + k * F1_DT * vd[0], c0[1] + k * F1_DT * vd[1], 0.15] for k in range(N + 1)])
    vel = np.tile(vd, (N + 1, 1))
    am = np.zeros((N + 1, 3))
    des_state = np.concatenate([pos.ravel(), vel.ravel(), am.ravel()])
    table = np.array([[1, 0, 1, 0]] * 5 + [[0, 1, 0, 1]] * 5, dtype=np.float64)
    feet = np.repeat(F1_STATE[9:21].reshape(4, 1, 3), N + 1, axis=1)
    return cfg, F1_STATE.copy(), des_state, pack_des_inputs(table.T, feet)


def pack_des_inputs(contact, des_feet):
    """contact [L, N]; des_feet [L, N+1, 3] -> flat des_inputs (CentroidalMPC.cpp:316-317)."""
    L, N = contact.shape
    out = np.zeros(L * (4 * N + 3))
    for i in range(L):
        o = i * (4 * N + 3)
        out[o:o + N] = contact[i]
        out[o + N:o + 4 * N + 3] = np.asarray(des_feet[i], dtype=np.float64).reshape(-1)
    return out


def gait_table(gait, N, phase, L=4):
    """Contact table [L, N] of 0/1 doubles.  Leg order lf, rf, rh, lh
    (CentoidMPCTest.cpp:43-46).  Period 10 steps.  Every column has >=1 stance leg
    (a zero-stance step is invalid in the reference, CentroidalMPC.cpp:328-330).
    trot = diagonal pairs {0,2}/{1,3}; pace = same-side pairs {0,3}/{1,2}
    (gait.info LF_LH / RF_RH pattern without the flight gaps); bound = front {0,1} /
    hind {2,3}; gallop = rotary 4-beat with overlapping stance windows."""
    t = np.zeros((L, N))
    for j in range(N):
        ph = (j + phase) % 10
        first = ph < 5
        for i in range(L):
            if gait == "stand":
                on = True
            elif gait == "trot":
                on = first == (i in (0, 2))
            elif gait == "pace":
                on = first == (i in (0, 3))
            elif gait == "bound":
                on = first == (i in (0, 1))
            elif gait == "gallop":
                # leg i touches down at 2.5*order[i] and stays 4 steps -> 1-2 legs in stance
                order = {0: 0, 1: 1, 2: 2, 3: 3}[i]
                start = (order * 5) // 2
                on = ((ph - start) % 10) < 4
            else:
                raise ValueError(gait)
            t[i, j] = 1.0 if on else 0.0
    assert np.all(t.sum(axis=0) > 0)
    return t


def hard_config(N=10, mu=0.3, dt=0.03, wf=1e-2, disc_mode=0):
    """The CONSTRAINED workload: tracking-heavy weights (state tracking dominates force tracking) + low friction, so
    that friction-pyramid rows go active in (almost) every instance and the presolve cannot settle them.  Not a
    BASELINE.json config -- the reference driver's weights never activate a row -- but the workload every
    "identical active set" claim and every constrained-path throughput number of this repo is made on."""
    w = np.array([5e4, 5e4, 300, 500, 500, 500, 200, 200, 200] + [0.2] * 12 + [wf] * 12 + [wf / 10] * 12)
    return default_config(N, dt=dt, mu=[mu] * 4, weights=w, disc_mode=disc_mode)


def make_batch(cfg, B, first=0, gaits=("trot",), hard_fraction=0.25, seed=SEED):
    """Instances [first, first+B) of the synthetic workload.

    Returns state [B, 9+3L], des_state [B, 9(N+1)], des_inputs [B, L(4N+3)].
    'hard' instances (every 4th by default) ask for a ~1 m/s velocity change so that
    friction-pyramid rows go active.  NOTE: mu is part of the (shared) config in the
    reference ctor, so low-friction cases are separate *configs*, not per instance.
    """
    N, L = cfg["horizon"], cfg["num_legs"]
    assert L == 4
    sz = sizes(cfg)
    st = np.zeros((B, sz["state"]))
    ds = np.zeros((B, sz["des_state"]))
    di = np.zeros((B, sz["des_inputs"]))
    feet0 = F1_STATE[9:21].reshape(4, 3)
    dt = cfg["dt"]
    period = max(1, int(round(1.0 / max(hard_fraction, 1e-9)))) if hard_fraction > 0 else 0
    for b in range(B):
        idx = first + b
        rng = np.random.Generator(np.random.Philox(key=seed, counter=[idx, 0, 0, 0]))
        s = int(rng.integers(0, 10))
        c = np.array([0, 0, 0.15]) + rng.uniform(-0.02, 0.02, 3)
        v = np.concatenate([rng.uniform(-0.3, 0.3, 2), rng.uniform(-0.1, 0.1, 1)])
        lm = rng.uniform(-0.1, 0.1, 3)
        feet = feet0 + np.concatenate([rng.uniform(-0.01, 0.01, (4, 2)), np.zeros((4, 1))], axis=1)
        vd = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), 0.0])
        ang = rng.uniform(0, 2 * np.pi)
        if period and idx % period == period - 1:
            vd[:2] = v[:2] + 1.0 * np.array([np.cos(ang), np.sin(ang)])
        gait = gaits[idx % len(gaits)]
        st[b, 0:3], st[b, 3:6], st[b, 6:9] = c, v, lm
        st[b, 9:] = feet.ravel()
        pos = np.array([[c[0] + k * dt * vd[0], c[1] + k * dt * vd[1], 0.15] for k in range(N + 1)])
        ds[b] = np.concatenate([pos.ravel(), np.tile(vd, N + 1), np.zeros(3 * (N + 1))])
        di[b] = pack_des_inputs(gait_table(gait, N, s), np.repeat(feet[:, None, :], N + 1, axis=1))
    return st, ds, di
