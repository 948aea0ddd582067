"""-m gpu: the C++ shim (reference class surface) and the closed-loop rollout (BASELINE config 5)."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_shim_reproduces_reference_driver(pkg, orc, wl):
    """tests/cpp/centoid_mpc_test.cpp = CentoidMPCTest.cpp against cheeta-mpc_b200/include/CentroidalMPC.h."""
    exe = os.path.join(ROOT, "cheeta-mpc_b200", "csrc", "centoid_mpc_test")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][0]
    out = json.loads(line)
    assert out["invalid_table_throws"] is True
    cfg, st, ds, di = wl.fixture_f1()
    ref = orc.solve_batch(pkg.make_config(cfg), st[None], ds[None], di[None])
    f = np.array(out["forces"])
    assert np.abs(f - ref["forces"][0]).max() <= 1e-6 * np.abs(ref["forces"][0]).max()
    assert "finished test" in r.stdout


def advance_reference(orc, ccfg, cfg, st, ds, di, hip, forces):
    """NumPy statement of one closed-loop tick (include/cmpc.h, cmpc_rollout): reference plant
    (oracle plant_step), table rotation, swing feet carried under the hips, references
    regenerated from the new state."""
    N, L, dt = cfg["horizon"], cfg["num_legs"], cfg["dt"]
    B = len(st)
    st, ds, di = st.copy(), ds.copy(), di.copy()
    for b in range(B):
        D = di[b].reshape(L, 4 * N + 3)
        f0 = forces[b].reshape(L, N, 3)[:, 0, :]
        st[b, :9] = orc.plant_step(ccfg, st[b, :9], st[b, 9:].reshape(L, 3), D[:, 0].copy(), f0)
        x = st[b]
        S = ds[b].reshape(3, N + 1, 3)
        vd, zd, ad = S[1, 0].copy(), S[0, 0, 2], S[2, 0].copy()
        for k in range(N + 1):
            S[0, k] = [x[0] + k * dt * vd[0], x[1] + k * dt * vd[1], zd]
            S[1, k] = vd
            S[2, k] = ad
        D[:, :N] = np.roll(D[:, :N], -1, axis=1)
        feet = D[:, N:].reshape(L, N + 1, 3)
        for i in range(L):
            if not D[i, 0] > 0:
                x[9 + 3 * i:12 + 3 * i] = [x[0] + hip[b, i, 0], x[1] + hip[b, i, 1], hip[b, i, 2]]
            planted = True
            for k in range(N + 1):
                planted = planted and D[i, min(k, N - 1)] > 0
                feet[i, k] = x[9 + 3 * i:12 + 3 * i] if planted else [S[0, k, 0] + hip[b, i, 0], S[0, k, 1] + hip[b, i, 1], hip[b, i, 2]]
    return st, ds, di


@pytest.mark.parametrize("warm,presolve", [(0, 1), (1, 1), (0, 0), (1, 0)])
def test_rollout_matches_tickwise_oracle(pkg, orc, wl, warm, presolve):
    from conftest import hard_config
    for cfg in (dict(wl.default_config(10), presolve=presolve), dict(hard_config(wl, 10, 0.3), presolve=presolve)):
        B, ticks = 40, 14
        st, ds, di = wl.make_batch(cfg, B, gaits=wl.GAITS)
        m = pkg.CentroidalMPC.from_dict(cfg)
        m.SetupMPC(B)
        out = m.Rollout(st, ds, di, ticks, warm_start=warm)
        rs, rds, rdi = st.copy(), ds.copy(), di.copy()
        hip = rs[:, 9:].reshape(B, 4, 3) - np.concatenate([rs[:, :2], np.zeros((B, 1))], axis=1)[:, None, :]
        cold_iters = np.zeros(B, np.int64)
        for t in range(ticks):
            ref = orc.solve_batch(m.cfg, rs, rds, rdi, nthreads=8, want_lam=False)
            assert (ref["status"] == 0).all()
            cold_iters += ref["iters"]
            f0 = ref["forces"].reshape(B, 4, 10, 3)[:, :, 0, :].reshape(B, 12)
            assert np.abs(out["force_log"][t] - f0).max() <= 1e-6 * np.abs(f0).max(), (warm, t)
            rs, rds, rdi = advance_reference(orc, m.cfg, cfg, rs, rds, rdi, hip, ref["forces"])
        assert np.abs(out["state"] - rs).max() <= 1e-8 * (1 + np.abs(rs).max())
        assert np.array_equal(out["des_inputs"][:, :10], rdi[:, :10])
        assert np.abs(out["des_inputs"] - rdi).max() <= 1e-9 and np.abs(out["des_state"] - rds).max() <= 1e-9
        assert (out["status_or"] == 1).all()          # only CMPC_STATUS_OK seen
        if warm:   # verified guesses skip the interior-point iterations (nothing left to skip when the presolve settles every tick)
            assert out["iters_sum"].sum() < cold_iters.sum() or cold_iters.sum() == 0
        else:
            assert out["stats"]["launches"] >= 3 * ticks
        m.close()


def test_rollout_stays_physical(pkg, wl):
    """Closed loop sanity, 200 ticks of trot: with tracking-capable weights and moderate velocity
    commands the height holds near the desired 0.15 m and the commanded velocity is tracked; with
    the reference driver's weights (force tracking dominates, CentoidMPCTest.cpp:19-33) the loop
    is only required to stay finite and solved."""
    from conftest import hard_config
    cfg = hard_config(wl, 10, 0.8)
    st, ds, di = wl.make_batch(cfg, 64, hard_fraction=0.0)
    m = pkg.CentroidalMPC.from_dict(cfg)
    m.SetupMPC(64)
    out = m.Rollout(st, ds, di, 200, warm_start=1, log_forces=False)
    assert (out["status_or"] == 1).all()
    z, v, vd = out["state"][:, 2], out["state"][:, 3:5], ds[:, 33:35]
    print("height range", z.min(), z.max(), "velocity error", np.abs(v - vd).max())
    assert np.abs(z - 0.15).max() < 0.05
    assert np.abs(v - vd).max() < 0.5
    m.close()
    cfg = wl.default_config(10)
    st, ds, di = wl.make_batch(cfg, 64)
    m = pkg.CentroidalMPC.from_dict(cfg)
    m.SetupMPC(64)
    out = m.Rollout(st, ds, di, 200, warm_start=1, log_forces=False)
    assert (out["status_or"] == 1).all() and np.isfinite(out["state"]).all()
    m.close()


def test_solve_is_deterministic_and_batch_invariant(pkg, wl):
    """Size-independent properties at BASELINE's full batch: identical results when the same
    instance sits in a different batch position / batch size; run-to-run bit identical."""
    cfg = wl.default_config(10)
    B = 4096
    st, ds, di = wl.make_batch(cfg, B)
    m = pkg.CentroidalMPC.from_dict(cfg)
    m.SetupMPC(B)
    a = m.UpdateMPCBatch(st, ds, di, want_lam=False)
    b = m.UpdateMPCBatch(st, ds, di, want_lam=False)
    assert np.array_equal(a["forces"], b["forces"])
    assert (a["status"] == 0).all() and a["kkt"].max() <= 1e-8
    perm = np.random.default_rng(0).permutation(B)[:512]
    c = m.UpdateMPCBatch(st[perm], ds[perm], di[perm], want_lam=False)
    assert np.array_equal(c["forces"], a["forces"][perm])
    # physics sanity on every instance: swing legs carry no force, stance fz inside the bounds
    F = a["forces"].reshape(B, 4, 10, 3)
    contact = di.reshape(B, 4, 43)[:, :, :10]
    assert np.all(F[contact == 0] == 0)
    fz = F[..., 2][contact > 0]
    assert fz.min() > 0 and fz.max() <= 8 * 9.81 * 4
    m.close()


def test_weights_update_and_zoh_mode(pkg, orc, wl):
    from conftest import hard_config
    cfg = wl.default_config(10)
    st, ds, di = wl.make_batch(cfg, 32, gaits=wl.GAITS)
    m = pkg.CentroidalMPC.from_dict(cfg)
    m.SetupMPC(32)
    hard = hard_config(wl, 10, 0.8)
    m.UpdateWeights(hard["weights"])                     # NonlinearMPC::UpdateWeights
    out = m.UpdateMPCBatch(st, ds, di)
    cfg2 = dict(cfg, weights=hard["weights"])
    ref = orc.solve_batch(pkg.make_config(cfg2), st, ds, di)
    assert np.abs(out["forces"] - ref["forces"]).max() <= 1e-6 * np.abs(ref["forces"]).max()
    with pytest.raises(pkg.CmpcError):
        m.UpdateWeights(hard["weights"][:10])
    m.close()
    cfgz = hard_config(wl, 10, 0.3, disc_mode=1)
    stz, dsz, diz = wl.make_batch(cfgz, 32, gaits=wl.GAITS)
    mz = pkg.CentroidalMPC.from_dict(cfgz)
    mz.SetupMPC(32)
    out = mz.UpdateMPCBatch(stz, dsz, diz)
    ref = orc.solve_batch(mz.cfg, stz, dsz, diz)
    assert (out["status"] == ref["status"]).all()
    assert np.abs(out["forces"] - ref["forces"]).max() <= 1e-6 * np.abs(ref["forces"]).max()
    assert np.array_equal(out["active"], ref["active"])
    mz.close()


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])
def test_host_buffer_paths_agree(pkg, wl, mode, monkeypatch):
    """cmpc_solve_batch with host buffers: mode 0 = automatic (pipelined chunks for big batches, else
    zero-copy), 1 = zero-copy reads of pinned inputs (outputs written in place when pinned), 2 = copy in /
    compute / copy out, 3 = progressive (chunked DMA copy-in overlapped with the router kernel that polls
    for its chunk), 4 = pipelined (chunk c computes while chunk c + 1 is copied), 5 = full duplex (pipelined,
    and each chunk's outputs leave by the copy engine on a third stream).  Pageable and pinned buffers, every mode: bit-identical results (mixed gaits, so the router also forwards the stand
    instances to their size class)."""
    import ctypes as C
    import torch
    monkeypatch.setenv("CMPC_E2E_MODE", str(mode))
    cfg = wl.default_config(10)
    B = 3000                                                    # ragged last chunk
    st, ds, di = wl.make_batch(cfg, B, gaits=wl.GAITS)
    m = pkg.CentroidalMPC.from_dict(cfg)
    m.SetupMPC(B)
    pageable = m.UpdateMPCBatch(st, ds, di)                     # pageable numpy buffers
    assert (pageable["status"] == 0).all()
    pin = [torch.from_numpy(a).pin_memory() for a in (st, ds, di)]
    f = torch.zeros(B, m.n_forces, dtype=torch.float64).pin_memory()
    s = torch.full((B,), -1, dtype=torch.int32).pin_memory()
    it = torch.zeros(B, dtype=torch.int32).pin_memory()
    kk = torch.zeros(B, dtype=torch.float64).pin_memory()
    act = torch.zeros(B, 10, 4, dtype=torch.int16).pin_memory()
    stats = pkg.CmpcStats()
    vp = C.c_void_p
    for _ in range(3):                                          # repeated calls reuse the ready flag
        f.zero_(); s.fill_(-1)
        rc = m.lib.cmpc_solve_batch(m.h, B, vp(pin[0].data_ptr()), vp(pin[1].data_ptr()), vp(pin[2].data_ptr()), vp(f.data_ptr()),
                                    vp(s.data_ptr()), vp(it.data_ptr()), vp(kk.data_ptr()), None, vp(act.data_ptr()), C.byref(stats))
        assert rc == 0
        assert np.array_equal(f.numpy(), pageable["forces"])
        assert np.array_equal(s.numpy(), pageable["status"]) and np.array_equal(it.numpy(), pageable["iters"])
        assert np.array_equal(act.numpy().view(np.uint16), pageable["active"])
    if mode not in (2, 5):
        assert stats.d2h_ms < 0.05                              # pinned outputs are written in place
    route = m.last_route()
    assert isinstance(route, str) and route and route != "none"
    if mode in (4, 5):
        assert ("duplex" in route) == (mode == 5), route
    # device-resident reference
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(a).to(dev) for a in (st, ds, di)]
    df = torch.zeros(B, m.n_forces, dtype=torch.float64, device=dev)
    dst = torch.zeros(B, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()                                    # the handle's stream does not order behind torch's
    m.solve_device(B, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), df.data_ptr(), dst.data_ptr())
    m.synchronize()
    assert np.array_equal(df.cpu().numpy(), pageable["forces"])
    m.close()


def test_auto_route_tuning_keeps_results(pkg, wl):
    """Default mode with a big pinned batch: the first six calls alternate between the zero-copy and the
    pipelined route (the library keeps the faster one afterwards); every call must return the same bits."""
    import ctypes as C
    import torch
    cfg = wl.default_config(10)
    B = 4096
    st, ds, di = wl.make_batch(cfg, B)
    m = pkg.CentroidalMPC.from_dict(cfg)
    m.SetupMPC(B)
    pin = [torch.from_numpy(a).pin_memory() for a in (st, ds, di)]
    f = torch.zeros(B, m.n_forces, dtype=torch.float64).pin_memory()
    s = torch.full((B,), -1, dtype=torch.int32).pin_memory()
    vp = C.c_void_p
    ref = None
    for call in range(9):
        f.zero_(); s.fill_(-1)
        rc = m.lib.cmpc_solve_batch(m.h, B, vp(pin[0].data_ptr()), vp(pin[1].data_ptr()), vp(pin[2].data_ptr()), vp(f.data_ptr()),
                                    vp(s.data_ptr()), None, None, None, None, None)
        assert rc == 0 and (s.numpy() == 0).all()
        if ref is None:
            ref = f.numpy().copy()
        assert np.array_equal(f.numpy(), ref), call
    m.close()
