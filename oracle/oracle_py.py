"""ctypes view of the CPU oracle (oracle/cmpc_oracle.c) -- TEST INFRASTRUCTURE ONLY.
Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcmpc_oracle.so")
_LIB = None


def _cpu_sig():
    import hashlib
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return hashlib.md5((line.rstrip("\n") + "\n").encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build(force=False):
    """(Re)build when a source is newer than the library, or when the library was built on a different CPU
    (the baseline port is compiled with -march=native; built artefacts travel to the GPU box)."""
    srcs = [os.path.join(_HERE, f) for f in ("cmpc_oracle.c", "cmpc_cpu_fast.c", "cmpc_oracle.h", "Makefile")]
    sig = os.path.join(_HERE, "_build", "cpu_sig")
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(_SO) < os.path.getmtime(f) for f in srcs)
    if not stale:
        try:
            stale = open(sig).read().strip() != _cpu_sig()
        except OSError:
            stale = True
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        build()
        l = C.CDLL(_SO)
        vp = C.c_void_p
        l.cmpc_oracle_build.argtypes = [vp] * 7
        l.cmpc_oracle_solve.argtypes = [vp] * 10
        l.cmpc_oracle_solve_batch.argtypes = [vp, C.c_int] + [vp] * 9 + [C.c_int]
        l.cmpc_fast_solve_batch.argtypes = [vp, C.c_int] + [vp] * 8 + [C.c_int]
        l.cmpc_oracle_plant_step.argtypes = [vp] * 6
        l.cmpc_oracle_plant_step.restype = None
        _LIB = l
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build_qp(ccfg, state, des_state, des_inputs):
    """ccfg: CmpcConfig (ctypes). Returns H [p,p], g [p], status."""
    p = 3 * ccfg.num_legs * ccfg.horizon
    H = np.zeros((p, p)); g = np.zeros(p); st = C.c_int32(0)
    s, d, i = _f(state), _f(des_state), _f(des_inputs)
    lib().cmpc_oracle_build(C.addressof(ccfg), _p(s), _p(d), _p(i), _p(H), _p(g), C.addressof(st))
    return H, g, st.value


def solve_batch(ccfg, state, des_state, des_inputs, nthreads=1, want_lam=True):
    N, L = ccfg.horizon, ccfg.num_legs
    s = _f(state); B = s.shape[0] if s.ndim == 2 else 1
    s = s.reshape(B, -1); d = _f(des_state).reshape(B, -1); i = _f(des_inputs).reshape(B, -1)
    assert s.shape[1] == 9 + 3 * L and d.shape[1] == 9 * (N + 1) and i.shape[1] == L * (4 * N + 3)
    forces = np.zeros((B, 3 * L * N)); status = np.zeros(B, np.int32); iters = np.zeros(B, np.int32)
    kkt = np.zeros(B); lam = np.zeros((B, 2, N, L, 5)) if want_lam else None
    active = np.zeros((B, N, L), np.uint16)
    lib().cmpc_oracle_solve_batch(C.addressof(ccfg), B, _p(s), _p(d), _p(i), _p(forces), _p(status), _p(iters),
                                  _p(kkt), _p(lam), _p(active), int(nthreads))
    return dict(forces=forces, status=status, iters=iters, kkt=kkt, lam=lam, active=active)


def fast_solve_batch(ccfg, state, des_state, des_inputs, nthreads=1):
    """The CPU baseline port (cmpc_cpu_fast.c): same outputs as solve_batch, no multipliers."""
    N, L = ccfg.horizon, ccfg.num_legs
    s = _f(state); B = s.shape[0] if s.ndim == 2 else 1
    s = s.reshape(B, -1); d = _f(des_state).reshape(B, -1); i = _f(des_inputs).reshape(B, -1)
    forces = np.zeros((B, 3 * L * N)); status = np.zeros(B, np.int32); iters = np.zeros(B, np.int32)
    kkt = np.zeros(B); active = np.zeros((B, N, L), np.uint16)
    rc = lib().cmpc_fast_solve_batch(C.addressof(ccfg), B, _p(s), _p(d), _p(i), _p(forces), _p(status), _p(iters),
                                     _p(kkt), _p(active), int(nthreads))
    assert rc == 0
    return dict(forces=forces, status=status, iters=iters, kkt=kkt, active=active)


def plant_step(ccfg, x, feet, contact, forces):
    xn = np.zeros(9)
    x, feet, contact, forces = _f(x), _f(feet), _f(contact), _f(forces)
    lib().cmpc_oracle_plant_step(C.addressof(ccfg), _p(x), _p(feet), _p(contact), _p(forces), _p(xn))
    return xn
