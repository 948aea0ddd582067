"""gait template -> contact table. CPU: the NumPy restatement of the vendored OCS2
gait logic against hand-derived tables; GPU: the device kernel bit-exact against the restatement,
and the generated tables solved with parity."""
import numpy as np
import pytest

import numpy_mirror as nm


def test_mirror_trot_template_matches_gait_info_by_hand():
    # gait.info:30-43: LF_RH on [0, 0.35), RF_LH on [0.35, 0.70); dt = 0.05 -> 7 steps per phase
    modes, times = nm.GAIT_INFO["trot"]
    t = nm.gait_contact_table(modes, times, 0.0, 0.05, 14)
    lf, rf, rh, lh = t
    assert lf[:7].tolist() == [1] * 7 and lf[7:].tolist() == [0] * 7       # LF_RH = {LF, RH}
    assert np.array_equal(lf, rh) and np.array_equal(rf, lh) and np.array_equal(lf + rf, np.ones(14))
    # wrapPhase: one period later / earlier gives the same table (Gait.cpp:63-69)
    assert np.array_equal(nm.gait_contact_table(modes, times, 0.70, 0.05, 14), t)
    assert np.array_equal(nm.gait_contact_table(modes, times, -0.70, 0.05, 14), t)


def test_mirror_mode_numbers_and_upper_bound_edges():
    # stanceLeg2ModeNumber = RH + 2 LH + 4 RF + 8 LF (MotionPhaseDefinition.h:129-132)
    assert nm.MODE["LF_RH"] == 9 and nm.MODE["RF_LH"] == 6 and nm.MODE["STANCE"] == 15 and nm.MODE["LF_RF_LH"] == 14
    # a phase exactly on an event belongs to the NEXT mode (std::upper_bound, Gait.cpp:74-79)
    modes, times = ["LF", "RF", "LH", "RH"], [0.0, 0.25, 0.5, 0.75, 1.0]
    t = nm.gait_contact_table(modes, times, 0.0, 0.25, 4)
    assert t.tolist() == [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]]  # rows lf, rf, rh, lh
    # static_walk: always three stance legs; flying_trot has flight columns (invalid for this MPC)
    sw = nm.gait_contact_table(*nm.GAIT_INFO["static_walk"], 0.13, 0.01, 120)
    assert (sw.sum(axis=0) == 3).all()
    ft = nm.gait_contact_table(*nm.GAIT_INFO["flying_trot"], 0.0, 0.01, 60)
    assert (ft.sum(axis=0) == 0).any()


@pytest.mark.gpu
def test_device_gait_tables_bit_exact_and_solvable(pkg, orc, wl):
    names = ["stance", "trot", "standing_trot", "static_walk", "dynamic_walk", "amble", "standing_pace", "flying_trot"]
    gaits = [pkg.make_gait([nm.MODE[m] for m in nm.GAIT_INFO[n][0]], nm.GAIT_INFO[n][1]) for n in names]
    for N, dt in ((10, 0.01), (30, 0.015), (6, 0.05)):
        cfg = wl.default_config(N, dt=dt)
        B = 256
        st, ds, di = wl.make_batch(cfg, B)
        rng = np.random.default_rng(N)
        gid = rng.integers(0, len(names), B).astype(np.int32)
        t0 = rng.uniform(-2.0, 5.0, B)
        t0[:8] = [0.0, 0.35, 0.7, 0.3, 0.65, 0.27, 0.6, 1.2]           # on the switching times
        m = pkg.CentroidalMPC.from_dict(cfg)
        m.SetupMPC(B)
        out = m.FillContactTables(gaits, gid, t0, di)
        exp = di.copy()
        for b in range(B):
            tab = nm.gait_contact_table(*nm.GAIT_INFO[names[gid[b]]], t0[b], dt, N)
            exp[b].reshape(4, 4 * N + 3)[:, :N] = tab
        assert np.array_equal(out, exp)                                 # flags AND untouched foot positions
        # the generated tables go straight into the solve; flight columns are flagged, not fatal
        res = m.UpdateMPCBatch(st, ds, out, want_lam=False)
        ref = orc.solve_batch(m.cfg, st, ds, out, nthreads=8, want_lam=False)
        assert np.array_equal(res["status"], ref["status"])
        flight = np.array([(out[b].reshape(4, 4 * N + 3)[:, :N].sum(axis=0) == 0).any() for b in range(B)])
        assert np.array_equal(res["status"] == 3, flight) and flight.any()
        ok = res["status"] == 0
        assert ok.sum() > B // 2
        scale = np.abs(ref["forces"][ok]).max(axis=1, keepdims=True)
        assert (np.abs(res["forces"][ok] - ref["forces"][ok]) / scale).max() <= 1e-6
        m.close()


@pytest.mark.gpu
def test_gait_argument_validation(pkg, wl):
    cfg = wl.default_config(10)
    st, ds, di = wl.make_batch(cfg, 4)
    m = pkg.CentroidalMPC.from_dict(cfg)
    m.SetupMPC(4)
    bad = pkg.make_gait([9, 6], [0.0, 0.35, 0.30])                      # switching times must ascend
    with pytest.raises(pkg.CmpcError, match="ascend"):
        m.FillContactTables([bad], np.zeros(4, np.int32), np.zeros(4), di)
    with pytest.raises(pkg.CmpcError, match="0..15"):
        m.FillContactTables([pkg.make_gait([16], [0.0, 0.5])], np.zeros(4, np.int32), np.zeros(4), di)
    m.close()


# ---------------------------------------------------------------- foot plan (reference output foot_pos)
def _foot_case(wl, seed, N=10):
    cfg = wl.default_config(N)
    rng = np.random.default_rng(seed)
    st, ds, di = wl.make_batch(cfg, 1, first=seed, gaits=(wl.GAITS[seed % 5],))
    D = di[0].reshape(4, 4 * N + 3)
    D[:, N:] += rng.normal(scale=0.08, size=(4, 3 * (N + 1)))        # moving desired footholds
    return cfg, st[0], di[0]


def test_foot_plan_is_feasible_and_locally_optimal(wl):
    for seed in range(12):
        cfg, st, di = _foot_case(wl, seed)
        p = nm.foot_plan(cfg, st, di)
        cost, ok = nm.foot_cost_and_feasible(cfg, st, di, p)
        if not ok:
            continue  # current foot too far from a desired foothold in its initial stance run: the reference NLP is infeasible too
        rng = np.random.default_rng(100 + seed)
        D = di.reshape(4, 43)
        for _ in range(200):  # feasible perturbations never do better (convex problem: local = global)
            q = p.copy()
            i, k = rng.integers(0, 4), rng.integers(1, 11)
            a = k
            while a > 0 and 1.0 - D[i, a - 1] == 0.0:
                a -= 1
            e = k
            while e < 10 and 1.0 - D[i, e] == 0.0:
                e += 1
            if a == 0:
                continue
            q[i, a:e + 1] += rng.normal(scale=0.01, size=3)
            c2, ok2 = nm.foot_cost_and_feasible(cfg, st, di, q)
            assert (not ok2) or c2 >= cost - 1e-12


@pytest.mark.gpu
def test_device_foot_plan_matches_mirror(pkg, wl):
    cfg = wl.default_config(10)
    B = 128
    st, ds, di = wl.make_batch(cfg, B, gaits=wl.GAITS)
    rng = np.random.default_rng(5)
    di.reshape(B, 4, 43)[:, :, 10:] += rng.normal(scale=0.08, size=(B, 4, 33))
    m = pkg.CentroidalMPC.from_dict(cfg)
    m.SetupMPC(B)
    fp = m.FootPlan(st, di)
    for b in range(B):
        assert np.abs(fp[b] - nm.foot_plan(cfg, st[b], di[b])).max() <= 1e-15
    m.close()


def test_mirror_gait_switch_inserts_intermediate_stance_by_hand():
    """GaitSchedule::insertModeSequenceTemplate (GaitSchedule.cpp:47-72) by hand, on a time grid that is exact in binary
    (dt = 1/16): trot-like template [0, 0.375, 0.75] tiled from 0, switch to a pace-like template [0, 0.25, 0.5] at
    t = 0.5 with 0.25 s of intermediate stance."""
    trot = (["LF_RH", "RF_LH"], [0.0, 0.375, 0.75])
    pace = (["LF_LH", "RF_RH"], [0.0, 0.25, 0.5])
    dt = 0.0625
    t = nm.gait_switch_contact_table(trot, pace, 0.0, 0.5, 0.25, 0.0, dt, 20)
    # lower_bound: a step ON an event keeps the earlier mode.  j = 0 (t = 0 = the first event): the initial STANCE;
    # j = 1..6 (.. 0.375): LF_RH; j = 7, 8 (0.4375, 0.5): RF_LH runs on until t_switch; j = 9..12 (.. 0.75): the inserted
    # STANCE; the new template is tiled from 0.75: LF_LH on (0.75, 1.0], RF_RH on (1.0, 1.25]
    exp_modes = [15] + [9] * 6 + [6] * 2 + [15] * 4 + [nm.MODE["LF_LH"]] * 4 + [nm.MODE["RF_RH"]] * 3
    bit = [8, 4, 1, 2]
    for i in range(4):
        assert t[i].tolist() == [1.0 if m & bit[i] else 0.0 for m in exp_modes]
    # switching while the running mode already is STANCE inserts no extra stance: stance template -> trot at 0.25
    stance = (["STANCE"], [0.0, 0.5])
    t2 = nm.gait_switch_contact_table(stance, trot, 0.0, 0.25, 0.25, 0.0, dt, 12)
    exp2 = [15] * 5 + [9] * 6 + [6]
    for i in range(4):
        assert t2[i].tolist() == [1.0 if m & bit[i] else 0.0 for m in exp2]
    # no switch: the plain tiling; a tiling that starts after the horizon is all initial stance
    t3 = nm.gait_switch_contact_table(trot, pace, 0.0, None, 0.25, 0.0, dt, 16)
    assert t3[0].tolist() == [1.0] + [1.0] * 6 + [0.0] * 6 + [1.0] * 3
    assert nm.gait_switch_contact_table(trot, pace, 5.0, None, 0.25, 0.0, dt, 10).min() == 1.0


@pytest.mark.gpu
def test_device_gait_switch_bit_exact(pkg, orc, wl):
    """cmpc_fill_contact_tables_switch against the explicit schedule construction of the mirror, bit for bit: random template
    pairs, tiling starts, switch times (incl. steps that land exactly on events, switches before the tiling start, after
    the horizon, NaN = none), stance times 0 / 0.15 / 0.4; the tables then solve with parity."""
    names = ["stance", "trot", "standing_trot", "static_walk", "dynamic_walk", "amble", "standing_pace"]
    gaits = [pkg.make_gait([nm.MODE[m] for m in nm.GAIT_INFO[n][0]], nm.GAIT_INFO[n][1]) for n in names]
    for N, dt, Tst in ((10, 0.05, 0.4), (30, 0.02, 0.15), (16, 0.05, 0.0)):
        cfg = wl.default_config(N, dt=dt)
        B = 384
        st, ds, di = wl.make_batch(cfg, B)
        rng = np.random.default_rng(7 * N)
        ga = rng.integers(0, len(names), B).astype(np.int32)
        gb = rng.integers(0, len(names), B).astype(np.int32)
        t0 = np.round(rng.uniform(0.0, 3.0, B) / dt) * dt            # steps on a dt grid: many land exactly on events
        tt = t0 - np.round(rng.uniform(0.0, 2.0, B) / 0.05) * 0.05
        ts = t0 + np.round(rng.uniform(-0.3, N * dt + 0.2, B) / 0.05) * 0.05
        ts[::7] = np.nan
        tt[5::11] = t0[5::11] + 0.3                                   # tiling starts inside the horizon
        m = pkg.CentroidalMPC.from_dict(cfg)
        m.SetupMPC(B)
        out = m.FillContactTablesSwitch(gaits, ga, gb, tt, ts, Tst, t0, di)
        exp = di.copy()
        for b in range(B):
            tab = nm.gait_switch_contact_table(nm.GAIT_INFO[names[ga[b]]], nm.GAIT_INFO[names[gb[b]]], tt[b],
                                               None if np.isnan(ts[b]) else ts[b], Tst, t0[b], dt, N)
            exp[b].reshape(4, 4 * N + 3)[:, :N] = tab
        bad = np.nonzero((out != exp).any(axis=1))[0]
        assert len(bad) == 0, (N, bad[:5], ga[bad[:5]], gb[bad[:5]], tt[bad[:5]], ts[bad[:5]], t0[bad[:5]])
        assert (exp.reshape(B, 4, 4 * N + 3)[:, :, :N].sum(axis=1) > 0).all()   # these templates have no flight phase
        res = m.UpdateMPCBatch(st, ds, out, want_lam=False)
        ref = orc.solve_batch(m.cfg, st, ds, out, nthreads=8, want_lam=False)
        assert np.array_equal(res["status"], ref["status"]) and (res["status"] == 0).all()
        sc = np.abs(ref["forces"]).max(axis=1, keepdims=True)
        assert (np.abs(res["forces"] - ref["forces"]) / sc).max() <= 1e-6
        m.close()
