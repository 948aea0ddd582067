"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.
Tolerances (BASELINE.json north_star): forces <= 1e-6 relative, KKT <= 1e-8, identical
active set; H/g build <= 1e-12 relative to max|H| (different arithmetic order)."""
import numpy as np
import pytest

from conftest import hard_config

pytestmark = pytest.mark.gpu

FORCE_RTOL = 1e-6
KKT_TOL = 1e-8
BUILD_RTOL = 1e-12


def _mpc(pkg, cfg, B):
    m = pkg.CentroidalMPC.from_dict(cfg)
    m.SetupMPC(B)
    return m


def _compare(pkg, orc, cfg, st, ds, di, tag):
    m = _mpc(pkg, cfg, len(st))
    out = m.UpdateMPCBatch(st, ds, di)
    ref = orc.solve_batch(m.cfg, st, ds, di, nthreads=8)
    assert (out["status"] == ref["status"]).all(), (tag, np.bincount(out["status"]), np.bincount(ref["status"]))
    scale = np.abs(ref["forces"]).max(axis=1, keepdims=True) + 1e-300
    err = (np.abs(out["forces"] - ref["forces"]) / scale).max()
    assert err <= FORCE_RTOL, (tag, err)
    ok = out["status"] <= 1
    assert out["kkt"][ok].max() <= KKT_TOL, (tag, out["kkt"][ok].max())
    assert (out["active"] == ref["active"]).all(), (tag, int((out["active"] != ref["active"]).sum()))
    m.close()
    return out, ref, err


@pytest.mark.parametrize("fx", ["fixture_f1", "fixture_f1_intended", "fixture_f1_n10"])
def test_fixture_parity(pkg, orc, wl, fx):
    cfg, st, ds, di = getattr(wl, fx)()
    out, ref, err = _compare(pkg, orc, cfg, st[None], ds[None], di[None], fx)
    assert out["status"][0] == 0


@pytest.mark.parametrize("N,disc", [(6, 0), (10, 0), (10, 1), (30, 0)])
def test_build_parity(pkg, orc, wl, N, disc):
    cfg = wl.default_config(N, disc_mode=disc)
    B = 10
    st, ds, di = wl.make_batch(cfg, B, gaits=wl.GAITS)
    m = _mpc(pkg, cfg, B)
    H, g, status = m.BuildQP(st, ds, di)
    for b in range(B):
        Ho, go, so = orc.build_qp(m.cfg, st[b], ds[b], di[b])
        assert status[b] == so
        assert np.abs(H[b] - Ho).max() <= BUILD_RTOL * np.abs(Ho).max()
        assert np.abs(g[b] - go).max() <= BUILD_RTOL * max(1.0, np.abs(go).max())
        assert np.array_equal(H[b], H[b].T)
    m.close()


def test_build_parity_hard(pkg, orc, wl):
    cfg = hard_config(wl, 10, 0.3)
    st, ds, di = wl.make_batch(cfg, 8, gaits=wl.GAITS)
    m = _mpc(pkg, cfg, 8)
    H, g, status = m.BuildQP(st, ds, di)
    for b in range(8):
        Ho, go, so = orc.build_qp(m.cfg, st[b], ds[b], di[b])
        assert np.abs(H[b] - Ho).max() <= BUILD_RTOL * np.abs(Ho).max()
        assert np.abs(g[b] - go).max() <= BUILD_RTOL * np.abs(go).max()
    m.close()


def test_headline_sample_parity(pkg, orc, wl):
    cfg = wl.default_config(10)
    st, ds, di = wl.make_batch(cfg, 512)
    out, ref, err = _compare(pkg, orc, cfg, st, ds, di, "config2")
    assert (out["status"] == 0).all()


def test_mixed_gaits_parity(pkg, orc, wl):
    cfg = wl.default_config(10)
    st, ds, di = wl.make_batch(cfg, 320, gaits=wl.GAITS)
    _compare(pkg, orc, cfg, st, ds, di, "config4")


@pytest.mark.parametrize("mu", [0.8, 0.3, 0.1])
def test_active_set_parity(pkg, orc, wl, mu):
    cfg = hard_config(wl, 10, mu)
    st, ds, di = wl.make_batch(cfg, 200, gaits=wl.GAITS)
    out, ref, err = _compare(pkg, orc, cfg, st, ds, di, f"hard mu={mu}")
    nact = sum(bin(int(a) & 0x3FF).count("1") for a in out["active"].ravel())
    assert nact > 0, "workload does not exercise the active set"


def test_horizon30_parity(pkg, orc, wl):
    cfg = wl.default_config(30)
    st, ds, di = wl.make_batch(cfg, 24, gaits=wl.GAITS)
    _compare(pkg, orc, cfg, st, ds, di, "config3")
    cfg = hard_config(wl, 30, 0.3)
    st, ds, di = wl.make_batch(cfg, 24, gaits=wl.GAITS)
    _compare(pkg, orc, cfg, st, ds, di, "config3 hard")


def test_invalid_and_nonfinite_instances_do_not_poison_batch(pkg, orc, wl):
    cfg = wl.default_config(10)
    st, ds, di = wl.make_batch(cfg, 16)
    di[3].reshape(4, 43)[:, 4] = 0.0          # step 4: no stance leg -> reference throws
    st[7, 2] = np.nan
    out, ref, err = _compare(pkg, orc, cfg, st, ds, di, "bad instances")
    assert out["status"][3] == 3 and out["status"][7] == 4
    assert (out["forces"][3] == 0).all() and (out["forces"][7] == 0).all()
    good = [b for b in range(16) if b not in (3, 7)]
    assert (out["status"][good] == 0).all()
    m = _mpc(pkg, cfg, 1)
    with pytest.raises(pkg.CmpcError, match="mpc table invalid"):
        m.UpdateMPC(st[3], ds[3], di[3])
    m.close()
