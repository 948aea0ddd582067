#!/usr/bin/env python
"""Secondary measurements for BASELINE.json configs 3, 4, 5 (not the headline line; bench.py is).
One JSON line per config. Device-resident timing with CUDA events around the library calls."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def timed_device(pkg, cfg, st, ds, di, steps=20, warmup=3):
    import torch
    B = len(st)
    mpc = pkg.CentroidalMPC.from_dict(cfg)
    mpc.SetupMPC(B)
    mpc.set_stream(torch.cuda.current_stream().cuda_stream)
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(a).to(dev) for a in (st, ds, di)]
    f = torch.zeros(B, mpc.n_forces, dtype=torch.float64, device=dev)
    s = torch.zeros(B, dtype=torch.int32, device=dev)
    it = torch.zeros(B, dtype=torch.int32, device=dev)
    k = torch.zeros(B, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    run = lambda: mpc.solve_device(B, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), f.data_ptr(), s.data_ptr(), it.data_ptr(), k.data_ptr())
    for _ in range(warmup):
        run()
    ms = []
    for _ in range(steps):
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    st_ = s.cpu().numpy(); its = it.cpu().numpy(); kk = k.cpu().numpy()
    mpc.close()
    return float(np.median(ms)), st_, its, kk


def main():
    pkg = ge.load_package()
    wl = pkg.workloads
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import hard_config
    out = []
    # config 4: mixed gaits (one GPU's share of 65536 at 4 GPUs)
    cfg = wl.default_config(10)
    B = 16384
    st, ds, di = wl.make_batch(cfg, B, gaits=wl.GAITS)
    ms, s, its, kk = timed_device(pkg, cfg, st, ds, di)
    out.append(dict(config="4: mixed gaits, N=10", batch=B, p50_ms=ms, solves_per_s=B / ms * 1e3, mean_iters=float(its.mean()),
                    status=np.bincount(s, minlength=5).tolist(), max_kkt=float(kk.max())))
    # tracking-heavy, low friction (active friction rows)
    cfg = hard_config(wl, 10, 0.3)
    B = 4096
    st, ds, di = wl.make_batch(cfg, B, gaits=wl.GAITS)
    ms, s, its, kk = timed_device(pkg, cfg, st, ds, di)
    out.append(dict(config="hard: tracking-heavy weights, mu=0.3, mixed gaits, N=10", batch=B, p50_ms=ms, solves_per_s=B / ms * 1e3,
                    mean_iters=float(its.mean()), status=np.bincount(s, minlength=5).tolist(), max_kkt=float(kk[s <= 1].max())))
    # config 3: horizon 30.  qp_backend 0 = automatic (stage-wise Riccati presolve above 20 free leg-steps),
    # 1 = condensed dense for every class (stand: n = 360, factor in L2, interior-point kernel only)
    for backend, tag in ((0, "Riccati presolve"), (1, "dense")):
        cfg = dict(wl.default_config(30), qp_backend=backend)
        for gait, B, n in (("trot", 1024, 180), ("stand", 1024 if backend == 0 else 256, 360)):
            st, ds, di = wl.make_batch(cfg, B, gaits=(gait,))
            ms, s, its, kk = timed_device(pkg, cfg, st, ds, di, steps=5 if B > 256 else 3, warmup=2)
            out.append(dict(config=f"3: N=30 {gait} (n={n}), {tag}", batch=B, p50_ms=ms, solves_per_s=B / ms * 1e3,
                            mean_iters=float(its.mean()), status=np.bincount(s, minlength=5).tolist(), max_kkt=float(kk.max())))
    cfg = dict(wl.default_config(30), qp_backend=0)
    B = 8192
    st, ds, di = wl.make_batch(cfg, B, gaits=wl.GAITS)
    ms, s, its, kk = timed_device(pkg, cfg, st, ds, di, steps=5, warmup=2)
    out.append(dict(config="3: N=30 mixed gaits, batch 8192, Riccati presolve", batch=B, p50_ms=ms, solves_per_s=B / ms * 1e3,
                    mean_iters=float(its.mean()), status=np.bincount(s, minlength=5).tolist(), max_kkt=float(kk.max())))
    # horizon 30 with active rows: Riccati presolve defers, dense interior-point kernel (n = 180..360) solves
    cfg = hard_config(wl, 30, 0.3)
    B = 512
    st, ds, di = wl.make_batch(cfg, B, gaits=wl.GAITS)
    ms, s, its, kk = timed_device(pkg, cfg, st, ds, di, steps=3, warmup=1)
    out.append(dict(config="hard N=30: tracking-heavy weights, mu=0.3, mixed gaits", batch=B, p50_ms=ms, solves_per_s=B / ms * 1e3,
                    mean_iters=float(its.mean()), status=np.bincount(s, minlength=5).tolist(), max_kkt=float(kk[s <= 1].max())))
    # config 5: closed loop
    cfg = wl.default_config(10)
    B, ticks = 4096, 1000
    st, ds, di = wl.make_batch(cfg, B)
    mpc = pkg.CentroidalMPC.from_dict(cfg)
    mpc.SetupMPC(B)
    mpc.Rollout(st, ds, di, 5, log_forces=False)
    for warm in (0, 1):
        t0 = time.perf_counter()
        r = mpc.Rollout(st, ds, di, ticks, warm_start=warm, log_forces=False)
        wall = time.perf_counter() - t0
        out.append(dict(config=f"5: closed loop {ticks} ticks x {B}, " + ("warm start (previous active set tried first)" if warm else "cold start each tick"),
                        batch=B, ticks=ticks, device_ms=r["stats"]["kernel_ms"], wall_s=wall,
                        ticks_per_s=ticks / (r["stats"]["kernel_ms"] * 1e-3), solves_per_s=B * ticks / (r["stats"]["kernel_ms"] * 1e-3),
                        mean_ipm_iters_per_tick=float(r["iters_sum"].mean() / ticks), status_or=np.bincount(r["status_or"]).tolist()))
    mpc.close()
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
