"""ctypes binding of include/cmpc.h and a Python mirror of the reference class surface
(``CentroidalMPC(mass, num_legs, horizon, dt, weights, mu)`` / ``SetupMPC()`` /
``UpdateMPC(state, des_state, des_inputs)``; reference CentroidalMPC.h:26-32).
Used by the tests and bench.py; the C++ shim in include/CentroidalMPC.h is the host-side
mirror a reference user would compile against."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

MAX_LEGS = 4
NUM_WEIGHTS = 9 + 9 * MAX_LEGS
STATUS_NAMES = {0: "OK", 1: "OK_IPM", 2: "MAX_ITER", 3: "INVALID_TABLE", 4: "NUMERICAL"}
_HERE = os.path.dirname(os.path.abspath(__file__))


class CmpcConfig(C.Structure):
    _fields_ = [("mass", C.c_double), ("num_legs", C.c_int32), ("horizon", C.c_int32),
                ("dt", C.c_double), ("mu", C.c_double * MAX_LEGS),
                ("weights", C.c_double * NUM_WEIGHTS), ("disc_mode", C.c_int32),
                ("max_iter", C.c_int32), ("ipm_tol", C.c_double), ("polish", C.c_int32),
                ("presolve", C.c_int32), ("qp_backend", C.c_int32), ("reserved", C.c_int32)]


class CmpcStats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
                ("mean_iters", C.c_double), ("max_iters", C.c_int32), ("n_ok", C.c_int32),
                ("n_ok_ipm", C.c_int32), ("n_max_iter", C.c_int32), ("n_invalid", C.c_int32),
                ("n_numerical", C.c_int32), ("max_kkt", C.c_double), ("launches", C.c_int32),
                ("reserved", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class CmpcGait(C.Structure):
    _fields_ = [("num_modes", C.c_int32), ("modes", C.c_int32 * 8), ("switching_times", C.c_double * 9)]


def make_gait(modes, switching_times) -> "CmpcGait":
    """modes: OCS2 ModeNumber values 0..15 (bits LF=8, RF=4, LH=2, RH=1); switching_times: len(modes)+1."""
    g = CmpcGait()
    g.num_modes = len(modes)
    for k, m in enumerate(modes):
        g.modes[k] = int(m)
    for k, t in enumerate(switching_times):
        g.switching_times[k] = float(t)
    return g


class CmpcError(RuntimeError):
    pass


def make_config(cfg: dict, max_iter=50, ipm_tol=1e-9, polish=1, presolve=1) -> CmpcConfig:
    """dict(mass, num_legs, horizon, dt, mu, weights[, disc_mode]) -> CmpcConfig
    (pure Python; the same defaults as cmpc_config_init)."""
    L = int(cfg["num_legs"])
    c = CmpcConfig()
    c.mass, c.num_legs, c.horizon, c.dt = float(cfg["mass"]), L, int(cfg["horizon"]), float(cfg["dt"])
    for i in range(L):
        c.mu[i] = float(cfg["mu"][i])
    w = list(cfg["weights"])
    assert len(w) >= 9 + 9 * L, "weights needs 9+9*num_legs entries"
    for i in range(9 + 9 * L):
        c.weights[i] = float(w[i])
    c.disc_mode = int(cfg.get("disc_mode", 0))
    c.max_iter, c.ipm_tol, c.polish = int(cfg.get("max_iter", max_iter)), float(cfg.get("ipm_tol", ipm_tol)), int(cfg.get("polish", polish))
    c.presolve = int(cfg.get("presolve", presolve))
    c.qp_backend = int(cfg.get("qp_backend", 0))
    return c


def lib_path() -> str:
    return os.path.join(_HERE, "csrc", "libcmpc_b200.so")


_LIB = None


def load_library():
    """Load libcmpc_b200.so (built in-tree by __graft_entry__.build()). Raises if missing:
    the product has no CPU fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise CmpcError(f"{p} not built -- run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(p)
    dp, ip, u16p = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint16)
    vp = C.c_void_p
    lib.cmpc_config_init.argtypes = [C.POINTER(CmpcConfig), C.c_double, C.c_int, C.c_int, C.c_double, dp, dp]
    lib.cmpc_create.argtypes = [C.POINTER(CmpcConfig), C.POINTER(vp)]
    lib.cmpc_setup.argtypes = [vp, C.c_int, C.c_int]
    lib.cmpc_update_weights.argtypes = [vp, dp, C.c_int]
    lib.cmpc_solve_batch.argtypes = [vp, C.c_int] + [vp] * 9 + [C.POINTER(CmpcStats)]
    lib.cmpc_solve_batch_device.argtypes = [vp, C.c_int] + [vp] * 9 + [C.POINTER(CmpcStats)]
    lib.cmpc_build_batch.argtypes = [vp, C.c_int] + [vp] * 6
    lib.cmpc_rollout.argtypes = [vp, C.c_int, C.c_int, C.c_int] + [vp] * 6 + [C.POINTER(CmpcStats)]
    lib.cmpc_stage_step_batch.argtypes = [vp, C.c_int, C.c_int] + [vp] * 8
    lib.cmpc_foot_plan_batch.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.cmpc_solve_batch_sqp.argtypes = [vp, C.c_int, C.c_int] + [vp] * 6
    lib.cmpc_fill_contact_tables.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp]
    lib.cmpc_fill_contact_tables_device.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp]
    lib.cmpc_fill_contact_tables_switch.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, C.c_double, vp, vp]
    lib.cmpc_fill_contact_tables_switch_device.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, C.c_double, vp, vp]
    lib.cmpc_set_stream.argtypes = [vp, vp]
    lib.cmpc_synchronize.argtypes = [vp]
    lib.cmpc_measure_fp64_peak.argtypes = [vp, dp]
    lib.cmpc_destroy.argtypes = [vp]
    lib.cmpc_destroy.restype = None
    lib.cmpc_last_error.argtypes = [vp]
    lib.cmpc_last_error.restype = C.c_char_p
    lib.cmpc_version.restype = C.c_char_p
    lib.cmpc_last_route.argtypes = [vp]
    lib.cmpc_last_route.restype = C.c_char_p
    _LIB = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class CentroidalMPC:
    """Python mirror of the reference class (CentroidalMPC.h:15-33), batched."""

    def __init__(self, mass, num_legs, predict_horizon, time_step, weights, mu, device=0, **knobs):
        self.lib = load_library()
        self.cfg = make_config(dict(mass=mass, num_legs=num_legs, horizon=predict_horizon,
                                    dt=time_step, weights=weights, mu=mu, **knobs))
        self.h = C.c_void_p()
        rc = self.lib.cmpc_create(C.byref(self.cfg), C.byref(self.h))
        if rc != 0:
            raise CmpcError(f"cmpc_create failed ({rc}): invalid constructor arguments")
        self.device = device
        self.max_batch = 0
        self.N, self.L = int(predict_horizon), int(num_legs)
        self.last_stats = CmpcStats()

    @classmethod
    def from_dict(cls, cfg, device=0, **knobs):
        kn = {k: cfg[k] for k in ("disc_mode", "max_iter", "ipm_tol", "polish", "presolve", "qp_backend") if k in cfg}
        kn.update(knobs)
        return cls(cfg["mass"], cfg["num_legs"], cfg["horizon"], cfg["dt"], cfg["weights"], cfg["mu"],
                   device=device, **kn)

    # sizes
    @property
    def n_state(self): return 9 + 3 * self.L
    @property
    def n_des_state(self): return 9 * (self.N + 1)
    @property
    def n_des_inputs(self): return self.L * (4 * self.N + 3)
    @property
    def n_forces(self): return 3 * self.L * self.N

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.cmpc_last_error(self.h)
            raise CmpcError(f"cmpc error {rc}: {msg.decode() if msg else ''}")

    def SetupMPC(self, max_batch=1):
        self._check(self.lib.cmpc_setup(self.h, int(max_batch), int(self.device)))
        self.max_batch = int(max_batch)

    def UpdateWeights(self, weights):
        w = _f64(weights)
        self._check(self.lib.cmpc_update_weights(self.h, w.ctypes.data_as(C.POINTER(C.c_double)), w.size))

    def UpdateMPC(self, state, des_state, des_inputs):
        """Single instance, reference signature. Returns forces [L*N*3] (per-leg order).
        Raises like the reference on an invalid table (CentroidalMPC.cpp:329)."""
        out = self.UpdateMPCBatch(_f64(state)[None], _f64(des_state)[None], _f64(des_inputs)[None])
        if out["status"][0] == 3:
            raise CmpcError("mpc table invalid")
        return out["forces"][0]

    def UpdateMPCBatch(self, state, des_state, des_inputs, want_lam=True):
        if self.max_batch == 0:
            self.SetupMPC(max(1, len(state)))
        st = _f64(state); B = st.shape[0]
        st = _f64(st, (B, self.n_state)); ds = _f64(des_state, (B, self.n_des_state)); di = _f64(des_inputs, (B, self.n_des_inputs))
        forces = np.zeros((B, self.n_forces)); status = np.zeros(B, np.int32); iters = np.zeros(B, np.int32)
        kkt = np.zeros(B); lam = np.zeros((B, 2, self.N, self.L, 5)) if want_lam else None
        active = np.zeros((B, self.N, self.L), np.uint16)
        stats = CmpcStats()
        self._check(self.lib.cmpc_solve_batch(self.h, B, _ptr(st), _ptr(ds), _ptr(di), _ptr(forces), _ptr(status),
                                              _ptr(iters), _ptr(kkt), _ptr(lam), _ptr(active), C.byref(stats)))
        self.last_stats = stats
        return dict(forces=forces, status=status, iters=iters, kkt=kkt, lam=lam, active=active, stats=stats.as_dict())

    def solve_device(self, B, d_state, d_des_state, d_des_inputs, d_forces, d_status, d_iters=0, d_kkt=0,
                     d_lam=0, d_active=0, stats=None):
        """Raw device pointers (ints). Asynchronous unless stats is given."""
        vp = C.c_void_p
        self._check(self.lib.cmpc_solve_batch_device(
            self.h, int(B), vp(d_state), vp(d_des_state), vp(d_des_inputs), vp(d_forces), vp(d_status),
            vp(d_iters or None), vp(d_kkt or None), vp(d_lam or None), vp(d_active or None),
            C.byref(stats) if stats is not None else None))

    def BuildQP(self, state, des_state, des_inputs):
        st = _f64(state); B = st.shape[0]
        st = _f64(st, (B, self.n_state)); ds = _f64(des_state, (B, self.n_des_state)); di = _f64(des_inputs, (B, self.n_des_inputs))
        p = self.n_forces
        H = np.zeros((B, p, p)); g = np.zeros((B, p)); status = np.zeros(B, np.int32)
        if self.max_batch == 0:
            self.SetupMPC(B)
        self._check(self.lib.cmpc_build_batch(self.h, B, _ptr(st), _ptr(ds), _ptr(di), _ptr(H), _ptr(g), _ptr(status)))
        return H, g, status

    def Rollout(self, state, des_state, des_inputs, ticks, warm_start=1, log_forces=True):
        st = _f64(state).copy(); B = st.shape[0]
        ds = _f64(des_state, (B, self.n_des_state)).copy(); di = _f64(des_inputs, (B, self.n_des_inputs)).copy()
        if self.max_batch == 0:
            self.SetupMPC(B)
        flog = np.zeros((ticks, B, 3 * self.L)) if log_forces else None
        iters = np.zeros(B, np.int32); stor = np.zeros(B, np.int32); stats = CmpcStats()
        self._check(self.lib.cmpc_rollout(self.h, B, int(ticks), int(warm_start), _ptr(st), _ptr(ds), _ptr(di),
                                          _ptr(flog), _ptr(iters), _ptr(stor), C.byref(stats)))
        return dict(state=st, des_state=ds, des_inputs=di, force_log=flog, iters_sum=iters, status_or=stor,
                    stats=stats.as_dict())

    def StageStep(self, state, des_state, des_inputs, hess, rhs, mode=1):
        """Stage-wise linear algebra alone (cmpc_stage_step_batch). Returns d_fused, d_resolve, grad [B, N*L*3]."""
        st = _f64(state); B = st.shape[0]
        st = _f64(st, (B, self.n_state)); ds = _f64(des_state, (B, self.n_des_state)); di = _f64(des_inputs, (B, self.n_des_inputs))
        hs = _f64(hess, (B, 6 * self.L * self.N)); r = _f64(rhs, (B, self.n_forces))
        if self.max_batch == 0:
            self.SetupMPC(B)
        o = [np.zeros((B, self.n_forces)) for _ in range(3)]
        self._check(self.lib.cmpc_stage_step_batch(self.h, B, int(mode), _ptr(st), _ptr(ds), _ptr(di), _ptr(hs), _ptr(r),
                                                   _ptr(o[0]), _ptr(o[1]), _ptr(o[2])))
        return o

    def FootPlan(self, state, des_inputs):
        """Optimal foot positions [B, L, N+1, 3] (the foot_pos outputs of the reference controller)."""
        st = _f64(state); B = st.shape[0]
        st = _f64(st, (B, self.n_state)); di = _f64(des_inputs, (B, self.n_des_inputs))
        if self.max_batch == 0:
            self.SetupMPC(B)
        fp = np.zeros((B, self.L, self.N + 1, 3))
        self._check(self.lib.cmpc_foot_plan_batch(self.h, B, _ptr(st), _ptr(di), _ptr(fp)))
        return fp

    def SolveSQP(self, state, des_state, des_inputs, sqp_iters=2):
        """Successive re-linearisation of the lever arms. Returns forces, status, defect [B, iters+1]."""
        st = _f64(state); B = st.shape[0]
        st = _f64(st, (B, self.n_state)); ds = _f64(des_state, (B, self.n_des_state)); di = _f64(des_inputs, (B, self.n_des_inputs))
        if self.max_batch == 0:
            self.SetupMPC(B)
        forces = np.zeros((B, self.n_forces)); status = np.zeros(B, np.int32); defect = np.zeros((B, sqp_iters + 1))
        self._check(self.lib.cmpc_solve_batch_sqp(self.h, B, int(sqp_iters), _ptr(st), _ptr(ds), _ptr(di), _ptr(forces), _ptr(status), _ptr(defect)))
        return dict(forces=forces, status=status, defect=defect)

    def FillContactTables(self, gaits, gait_id, t0, des_inputs):
        """Device-side gait -> contact table. gaits: list of CmpcGait; returns des_inputs copy."""
        di = _f64(des_inputs).copy(); B = di.shape[0]
        gid = np.ascontiguousarray(gait_id, dtype=np.int32); tt = _f64(t0)
        arr = (CmpcGait * len(gaits))(*gaits)
        if self.max_batch == 0:
            self.SetupMPC(B)
        self._check(self.lib.cmpc_fill_contact_tables(self.h, B, C.cast(arr, C.c_void_p), len(gaits), _ptr(gid), _ptr(tt), _ptr(di)))
        return di

    def FillContactTablesSwitch(self, gaits, gait_from, gait_to, t_tile, t_switch, stance_time, t0, des_inputs):
        """Gait switch with intermediate stance (GaitSchedule.cpp:47-72,107-137); returns a des_inputs copy."""
        di = _f64(des_inputs).copy(); B = di.shape[0]
        ga = np.ascontiguousarray(gait_from, dtype=np.int32); gb = np.ascontiguousarray(gait_to, dtype=np.int32)
        tt, ts, tz = _f64(t_tile), _f64(t_switch), _f64(t0)
        arr = (CmpcGait * len(gaits))(*gaits)
        if self.max_batch == 0:
            self.SetupMPC(B)
        self._check(self.lib.cmpc_fill_contact_tables_switch(self.h, B, C.cast(arr, C.c_void_p), len(gaits), _ptr(ga), _ptr(gb),
                                                             _ptr(tt), _ptr(ts), float(stance_time), _ptr(tz), _ptr(di)))
        return di

    def last_route(self):
        return self.lib.cmpc_last_route(self.h).decode()

    def set_stream(self, stream_ptr):
        self._check(self.lib.cmpc_set_stream(self.h, C.c_void_p(stream_ptr)))

    def synchronize(self):
        self._check(self.lib.cmpc_synchronize(self.h))

    def measure_fp64_peak(self):
        v = C.c_double()
        self._check(self.lib.cmpc_measure_fp64_peak(self.h, C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.cmpc_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
