#!/usr/bin/env python
"""Diagnostics for the stage-wise interior-point kernel (cmpc_ripm.cu): parity numbers against the oracle without
stopping at the first failure, then device-resident timings of the dense and the stage-wise route side by side."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as ge  # noqa: E402
from conftest import hard_config  # noqa: E402


def parity(pkg, orc, wl, cfg, B, gaits, tag):
    st, ds, di = wl.make_batch(cfg, B, gaits=gaits)
    m = pkg.CentroidalMPC.from_dict(cfg); m.SetupMPC(B)
    out = m.UpdateMPCBatch(st, ds, di)
    ref = orc.solve_batch(m.cfg, st, ds, di, nthreads=8)
    sc = np.abs(ref["forces"]).max(axis=1) + 1e-300
    err = np.abs(out["forces"] - ref["forces"]).max(axis=1) / sc
    ok = out["status"] <= 1
    print(json.dumps(dict(tag=tag, B=B, status=np.bincount(out["status"], minlength=5).tolist(),
                          ref_status=np.bincount(ref["status"], minlength=5).tolist(), max_force_err=float(err.max()),
                          n_bad=int((err > 1e-6).sum()), iters_mean=float(out["iters"].mean()), ref_iters_mean=float(ref["iters"].mean()),
                          iters_diff=int(np.abs(out["iters"] - ref["iters"]).max()), n_iter_diff=int((out["iters"] != ref["iters"]).sum()),
                          active_mismatch=int((out["active"] != ref["active"]).any(axis=(1, 2)).sum()),
                          max_kkt=float(out["kkt"][ok].max()) if ok.any() else None)), flush=True)
    m.close()


def stage(pkg, wl, N):
    import numpy_mirror as nm
    cfg = hard_config(wl, N, 0.3)
    B, L, nf = 4, 4, 12
    st, ds, di = wl.make_batch(cfg, B, gaits=wl.GAITS)
    rng = np.random.default_rng(N)
    sig = rng.uniform(0.01, 1e3, (B, N, L, 5)); rhs = rng.normal(size=(B, N * nf))
    hess = np.zeros((B, N, L, 6))
    for b in range(B):
        contact = nm.unpack(cfg, st[b], ds[b], di[b])[5]
        for k in range(N):
            for i in range(L):
                mu = cfg["mu"][i]; s = sig[b, k, i]
                sx, sy = s[0] + s[1], s[2] + s[3]
                hess[b, k, i] = [0.5 * sx, 0.5 * sy, 0.5 * (mu * mu * (sx + sy) + s[4]), 0.5 * mu * (s[1] - s[0]), 0.5 * mu * (s[3] - s[2]), 0]
                if not contact[i, k] > 0:
                    rhs[b, nf * k + 3 * i:nf * k + 3 * i + 3] = 0
    m = pkg.CentroidalMPC.from_dict(cfg); m.SetupMPC(B)
    d1, d2, g = m.StageStep(st, ds, di, hess, rhs, mode=1)
    for b in range(B):
        ref = nm.stage_newton_step(cfg, st[b], ds[b], di[b], sig[b], rhs[b])
        gref = nm.stage_gradient(cfg, st[b], ds[b], di[b], rhs[b])
        print(f"stage N={N} b={b}: fused {np.abs(d1[b]-ref).max()/np.abs(ref).max():.2e} resolve {np.abs(d2[b]-ref).max()/np.abs(ref).max():.2e} "
              f"grad {np.abs(g[b]-gref).max()/np.abs(gref).max():.2e}", flush=True)
    m.close()


def timed(pkg, cfg, st, ds, di, steps=10, warmup=3, default_stream=False):
    import torch
    B = len(st)
    mpc = pkg.CentroidalMPC.from_dict(cfg); mpc.SetupMPC(B)
    stream = torch.cuda.current_stream() if default_stream else torch.cuda.Stream()
    mpc.set_stream(stream.cuda_stream)
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(a).to(dev) for a in (st, ds, di)]
    f = torch.zeros(B, mpc.n_forces, dtype=torch.float64, device=dev)
    s = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
    k = torch.zeros(B, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    run = lambda: mpc.solve_device(B, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), f.data_ptr(), s.data_ptr(), it.data_ptr(), k.data_ptr())
    for _ in range(warmup):
        run()
    ms = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); run(); e1.record(stream); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    r = dict(ms=float(np.median(ms)), solves_per_s=B / float(np.median(ms)) * 1e3, iters=float(it.float().mean()),
             status=np.bincount(s.cpu().numpy(), minlength=5).tolist())
    mpc.close()
    return r


def main():
    pkg = ge.load_package(); orc = ge.load_oracle(); wl = pkg.workloads
    what = sys.argv[1:] or ["stage", "parity", "time"]
    if "stage" in what:
        for N in (6, 10, 30):
            stage(pkg, wl, N)
    if "parity" in what:
        for N, mu, B in ((10, 0.3, 256), (10, 0.1, 128), (6, 0.3, 64), (30, 0.3, 48)):
            parity(pkg, orc, wl, dict(hard_config(wl, N, mu), qp_backend=2), B, wl.GAITS, f"ripm hard N={N} mu={mu}")
        parity(pkg, orc, wl, dict(wl.default_config(10), qp_backend=2, presolve=0), 128, wl.GAITS, "ripm easy presolve=0")
    if "time" in what:
        for N, B, gaits in ((10, 4096, wl.GAITS), (10, 4096, ("trot",)), (10, 4096, ("stand",)), (30, 1024, wl.GAITS), (30, 4096, wl.GAITS)):
            cfg = hard_config(wl, N, 0.3)
            st, ds, di = wl.make_batch(cfg, B, gaits=gaits)
            for backend in (1, 2, 0):
                if backend == 1 and N == 30 and B > 1024:
                    continue
                r = timed(pkg, dict(cfg, qp_backend=backend), st, ds, di, steps=5 if N == 30 else 10)
                print(json.dumps(dict(config=f"hard mu=0.3 N={N} B={B} gaits={'/'.join(gaits)}", qp_backend=backend, **r)), flush=True)


if __name__ == "__main__":
    main()
