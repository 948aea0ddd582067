#!/usr/bin/env python
"""bench.py -- headline benchmark: MPC QP solves/sec (h=10, batch 4096 per GPU), p50 batch latency.

One "step" = one batched build+solve of `--batch` independent centroidal-MPC instances
(BASELINE.json config 2: horizon 10, trot gait, random initial states, reference weights)
through the C ABI.  `value` is timed with inputs resident in HBM (cmpc_solve_batch_device),
`e2e` through cmpc_solve_batch with pinned HOST buffers (H2D + kernel + D2H inside the
timed region).  N>1: one process per GPU (torchrun), each rank solves its own 4096
instances (weak scaling, no data-path collective), time = max over ranks.

`--impl reference` times the CPU restatement of the reference path (oracle/, all host
threads) on a bounded sample of the same workload: the reference's own CasADi/IPOPT build
is not buildable in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

METRIC = "MPC QP solves/sec (h=10, batch 4096 per GPU)"
UNIT = "solves/s"


def alg_flops_per_solve(n_free, iters, polish_dim):
    """SURVEY §8(d) useful-flop count, with p = number of FREE (stance) variables so that
    pinned swing-leg rows are not credited.  q = 9N rows are folded into closed forms, so
    the build is counted as the symmetric half of B'LB over free columns."""
    p = float(n_free)
    f_iter = p ** 3 / 3 + 4 * p ** 2
    f_polish = polish_dim ** 3 / 3 + 2 * polish_dim ** 2
    return iters * f_iter + f_polish


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region: NVML every 2 ms (nvidia_ml_py),
    falling back to the nvidia-smi query line of the profiling recipe (slower: a few samples per run)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False   # rows: (sm_mhz, sm_max_mhz, set(reasons))
        self.source = "nvml"

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            while not self.stop_flag:
                r = int(get(h))
                self.rows.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), float(mx),
                                  {k for k, b in bits.items() if r & b}))
                time.sleep(0.002)
            return
        except Exception:
            self.source = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                c = [x.strip() for x in out.strip().split(",")]
                if len(c) >= 7:
                    self.rows.append((float(c[0]), float(c[1]), {n for n, v in zip(names, c[3:7]) if v.lower().startswith("active")}))
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        sm = [r[0] for r in self.rows]
        reasons = set().union(*[r[2] for r in self.rows]) if self.rows else set()
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[1] for r in self.rows) if self.rows else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def cpu_baseline(pkg, cfg, st, ds, di, min_seconds=8.0):
    """CPU restatement (oracle/) on all host cores, bounded sample of the same workload."""
    orc = ge.load_oracle()
    ccfg = pkg.make_config(cfg)
    cores = os.cpu_count() or 1
    nsample = min(len(st), 4096)
    orc.solve_batch(ccfg, st[:64], ds[:64], di[:64], nthreads=cores, want_lam=False)  # warm-up
    done, t0 = 0, time.perf_counter()
    while True:
        orc.solve_batch(ccfg, st[:nsample], ds[:nsample], di[:nsample], nthreads=cores, want_lam=False)
        done += nsample
        el = time.perf_counter() - t0
        if el >= min_seconds:
            break
    return {"value": done / el, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{done} instances of the headline workload ({nsample}-instance batch x {done // nsample}), "
                      f"{el:.1f} s wall, one instance per pthread round-robin"}


def run_reference(args):
    """--impl reference: the reference path's CPU restatement, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = ge.load_package()
    wl = pkg.workloads
    cfg = dict(wl.default_config(horizon=args.horizon), presolve=args.presolve)
    nsample = 1024
    st, ds, di = wl.make_batch(cfg, nsample)
    orc = ge.load_oracle()
    ccfg = pkg.make_config(cfg)
    cores = os.cpu_count() or 1
    for _ in range(max(1, args.warmup)):
        orc.solve_batch(ccfg, st[:128], ds[:128], di[:128], nthreads=cores, want_lam=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.solve_batch(ccfg, st, ds, di, nthreads=cores, want_lam=False)
    el = time.perf_counter() - t0
    value = nsample * args.steps / el
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"config2 sample: {nsample} of batch {args.batch}, horizon {args.horizon}, trot, "
                                   "random initial states, reference weights (CentoidMPCTest.cpp:19-33)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{nsample} instances per step x {args.steps} steps; CPU restatement of the "
                                       "reference path (oracle/cmpc_oracle.c), NOT the reference's CasADi/IPOPT binary "
                                       "(not buildable here)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=4096, help="instances per GPU per step")
    ap.add_argument("--horizon", type=int, default=10)
    ap.add_argument("--gaits", default="trot")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--presolve", type=int, default=1, choices=[0, 1],
                    help="1 (library default): unconstrained minimiser tried first, IPM only for the rest; 0: IPM for all")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    pkg = ge.load_package()
    wl = pkg.workloads
    cfg = dict(wl.default_config(horizon=args.horizon), presolve=args.presolve)
    gaits = tuple(args.gaits.split(","))
    B = args.batch
    # rank r owns instance ids [r*B, (r+1)*B): contiguous block split, no inter-GPU traffic
    st, ds, di = wl.make_batch(cfg, B, first=rank * B, gaits=gaits)
    mpc = pkg.CentroidalMPC.from_dict(cfg, device=local_rank)
    mpc.SetupMPC(B)
    stream = torch.cuda.current_stream()
    mpc.set_stream(stream.cuda_stream)

    d_st, d_ds, d_di = (torch.from_numpy(a).to(dev) for a in (st, ds, di))
    d_forces = torch.zeros(B, mpc.n_forces, dtype=torch.float64, device=dev)
    d_status = torch.zeros(B, dtype=torch.int32, device=dev)
    d_iters = torch.zeros(B, dtype=torch.int32, device=dev)
    d_kkt = torch.zeros(B, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_device():
        mpc.solve_device(B, d_st.data_ptr(), d_ds.data_ptr(), d_di.data_ptr(), d_forces.data_ptr(),
                         d_status.data_ptr(), d_iters.data_ptr(), d_kkt.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step_device()
    torch.cuda.synchronize()
    st0 = pkg.CmpcStats()
    mpc.solve_device(B, d_st.data_ptr(), d_ds.data_ptr(), d_di.data_ptr(), d_forces.data_ptr(),
                     d_status.data_ptr(), d_iters.data_ptr(), d_kkt.data_ptr(), stats=st0)
    launches_per_step = int(st0.launches)   # classify + one solve kernel per size class
    status = d_status.cpu().numpy()
    iters = d_iters.cpu().numpy()
    kkt = d_kkt.cpu().numpy()
    forces0 = d_forces.cpu().numpy()
    assert (status <= 1).all(), f"unsolved instances: {np.bincount(status)}"

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s0, s1 in ev:
        flush.fill_(1)            # evict L2 between timed iterations (not timed)
        s0.record(stream)
        step_device()
        s1.record(stream)
    barrier()
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    total_ms = float(step_ms.sum())

    # ---- end to end through the public host-buffer call: pinned host -> H2D -> kernel -> D2H
    h_in = [torch.from_numpy(a).pin_memory() for a in (st, ds, di)]
    h_forces = torch.zeros(B, mpc.n_forces, dtype=torch.float64).pin_memory()
    h_status = torch.zeros(B, dtype=torch.int32).pin_memory()
    import ctypes as C
    lib, vp = mpc.lib, C.c_void_p

    def step_e2e():
        rc = lib.cmpc_solve_batch(mpc.h, B, vp(h_in[0].data_ptr()), vp(h_in[1].data_ptr()), vp(h_in[2].data_ptr()),
                                  vp(h_forces.data_ptr()), vp(h_status.data_ptr()), None, None, None, None, None)
        assert rc == 0, rc

    for _ in range(8):         # (the library times its two pinned-buffer routes during its first six calls)
        step_e2e()
    e2e_steps = max(10, min(args.steps, 100))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()             # synchronous: returns after the D2H copy completed
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    assert np.array_equal(h_forces.numpy(), forces0), "e2e path disagrees with device-resident path"
    sampler.stop_flag = True

    # ---- the same step with the presolve switched off (every instance through the interior-point
    # kernel): reported beside the headline so the two routes can be compared
    ipm_only = None
    if args.presolve and world == 1:
        m2 = pkg.CentroidalMPC.from_dict(dict(cfg, presolve=0), device=local_rank)
        m2.SetupMPC(B)
        m2.set_stream(stream.cuda_stream)

        def step2():
            m2.solve_device(B, d_st.data_ptr(), d_ds.data_ptr(), d_di.data_ptr(), d_forces.data_ptr(),
                            d_status.data_ptr(), d_iters.data_ptr(), d_kkt.data_ptr())
        for _ in range(3):
            step2()
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 50))]
        for s0, s1 in ev2:
            flush.fill_(1)
            s0.record(stream)
            step2()
            s1.record(stream)
        torch.cuda.synchronize()
        ms2 = float(np.median([a.elapsed_time(b) for a, b in ev2]))
        it2 = float(d_iters.cpu().numpy().mean())
        assert np.abs(d_forces.cpu().numpy() - forces0).max() <= 1e-6 * np.abs(forces0).max()
        ipm_only = {"value": B / (ms2 * 1e-3), "unit": UNIT, "p50_batch_latency_ms": ms2, "mean_ipm_iters": it2}
        m2.close()

    # ---- optional exchange (north star: "no inter-GPU traffic beyond an optional NCCL gather of
    # results over NVLink"): every rank ends up with all forces; timed on its own, NOT part of `value`
    gather = None
    if world > 1:
        allf = torch.empty(world * B, mpc.n_forces, dtype=torch.float64, device=dev)
        for _ in range(3):
            dist.all_gather_into_tensor(allf, d_forces)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        g0.record(stream)
        for _ in range(20):
            dist.all_gather_into_tensor(allf, d_forces)
        g1.record(stream)
        torch.cuda.synchronize()
        gms = torch.tensor([g0.elapsed_time(g1) / 20], dtype=torch.float64, device=dev)
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        assert torch.equal(allf[rank * B:(rank + 1) * B], d_forces)
        gather = {"collective": "ncclAllGather of the forces", "bytes_total": world * B * mpc.n_forces * 8,
                  "ms": float(gms[0]), "note": "outside the timed region; value is without it"}

    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_s_max = float(t[0]), float(t[1])

    if rank == 0:
        n_free = float(np.mean([(np.asarray(di[b]).reshape(4, -1)[:, :args.horizon] > 0).sum() * 3 for b in range(min(B, 256))]))
        mean_it = float(iters.mean())
        flops = alg_flops_per_solve(n_free, mean_it, n_free) * B
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        fp64_peak = mpc.measure_fp64_peak()
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if B == 4096 and args.horizon == 10 and gaits == ("trot",):
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        except Exception:
            pass
        med_ms = float(np.median(step_ms))
        achieved = flops / (med_ms * 1e-3) / 1e12
        in_bytes = 8 * (mpc.n_state + mpc.n_des_state + mpc.n_des_inputs)
        out_bytes = 8 * mpc.n_forces + 4 + 4 + 8
        hbm_gbs = B * (in_bytes + out_bytes) / (med_ms * 1e-3) / 1e9
        value = world * B * args.steps / (total_ms_max * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"config2: batch {B} per GPU, horizon {args.horizon}, gaits {','.join(gaits)}, random "
                                   "initial states seed 0xC0FFEE, reference weights, mu 0.8, dt 0.01, Euler",
                       "l2": "flushed between timed steps (256 MiB fill)", "sharding": f"dp{world} (independent instances)"},
            "p50_batch_latency_ms": med_ms, "p99_batch_latency_ms": float(np.percentile(step_ms, 99)),
            "mean_ipm_iters": mean_it, "max_kkt": float(kkt.max()),
            "status_counts": np.bincount(status, minlength=5).tolist(),
            "e2e": {"value": world * B * e2e_steps / e2e_s_max, "unit": UNIT,
                    "h2d_bytes_per_step": B * in_bytes, "d2h_bytes_per_step": B * (8 * mpc.n_forces + 4),
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s_max / e2e_steps},
            "gpu_launches": args.steps * launches_per_step,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak if fp64_peak else None, "traffic": traffic,
                         "peak_source": "DFMA microbenchmark run by this bench (MEASURED_PEAKS.json has no FP64 entry)",
                         "flops_per_solve": flops / B, "n_free": n_free,
                         "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": peaks.get("hbm_gbs"),
                                 "frac": hbm_gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                                 "bytes_per_solve": in_bytes + out_bytes}},
            "clocks": sampler.summary(),
        }
        if ipm_only is not None:
            f2 = alg_flops_per_solve(n_free, ipm_only["mean_ipm_iters"], n_free)
            ipm_only["roofline_frac"] = f2 * B / (ipm_only["p50_batch_latency_ms"] * 1e-3) / 1e12 / fp64_peak
            ipm_only["flops_per_solve"] = f2
            line["ipm_only"] = ipm_only
        line["config"]["presolve"] = args.presolve
        if gather is not None:
            line["nccl_gather"] = gather
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(pkg, cfg, st, ds, di)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    mpc.close()


if __name__ == "__main__":
    main()
