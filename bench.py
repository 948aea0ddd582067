#!/usr/bin/env python
"""bench.py -- headline benchmark: MPC QP solves/sec (h=10, batch 4096 per GPU), p50 batch latency.

One "step" = one batched build+solve of `--batch` independent centroidal-MPC instances
(BASELINE.json config 2: horizon 10, trot gait, random initial states, reference weights)
through the C ABI.  `value` is timed with inputs resident in HBM (cmpc_solve_batch_device),
`e2e` through cmpc_solve_batch with pinned HOST buffers (H2D + kernel + D2H inside the
timed region).  N>1: one process per GPU (torchrun), each rank solves its own `--batch`
instances (weak scaling, no data-path collective), time = max over ranks.

HEADLINE CAVEAT, stated in the line itself (`config.solver_path`): with the reference driver's weights no
friction row ever goes active, so the presolve (one Cholesky of H, verified) settles every instance and the
headline number is NOT a constrained-QP throughput.  The same JSON line therefore carries, as first-class blocks:
  constrained         tracking-heavy weights, mu 0.3, mixed gaits: every instance runs the interior-point + polish path
  config3 / config3_constrained   horizon 30 (BASELINE config 3) with the reference weights / with active rows
  config4             65 536 mixed-gait instances in total, split over the ranks (BASELINE config 4)
  strong              (N > 1) ONE 4096 batch split N ways -- what the BASELINE metric literally reads as
  ipm_only            the headline batch with the presolve off
Each block is timed like `value`: device-resident, CUDA events on the launching stream, L2 flushed between steps,
max over ranks.

`--impl reference` times the CPU BASELINE PORT of the reference path (oracle/cmpc_cpu_fast.c, all host threads,
persistent pool) on the SAME config and the SAME full batch per step: the reference's own CasADi/IPOPT build
cannot be compiled in this image.
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

UNIT = "solves/s"


def metric_name(args):
    return f"MPC QP solves/sec (h={args.horizon}, batch {args.batch} per GPU)"


def workload_name(args):
    tag = "config2" if (args.horizon == 10 and args.gaits == "trot") else ("config4 share" if args.horizon == 10 else "custom")
    return (f"{tag}: batch {args.batch} per GPU, horizon {args.horizon}, gaits {args.gaits}, random initial states seed "
            "0xC0FFEE, reference weights (CentoidMPCTest.cpp:19-33), mu 0.8, dt 0.01, Euler")


def alg_flops(n_free, iters, polish_dim=None):
    """Useful flops of one solve on n_free free (stance) variables: `iters` factorisations of H + C'SC with two
    solves each, plus one presolve / polish system.  Build, verification and vector work are not credited."""
    p = np.asarray(n_free, float)
    q = p if polish_dim is None else np.asarray(polish_dim, float)
    return iters * (p ** 3 / 3 + 4 * p ** 2) + (q ** 3 / 3 + 2 * q ** 2)


def stage_flops(m):
    """Useful flops of ONE stage of the stage-wise (Riccati) sweep with m free inputs over the augmented state
    (nz = 21): P Bbar exploiting the input structure (12 MACs per entry), the lower triangle of G, an m x m Cholesky,
    22 triangular solves, the symmetric half of P -= Y'Y, and the three vector sweeps of an iteration."""
    m = np.asarray(m, float)
    return 504 * m + 12 * m * (m + 1) + m ** 3 / 3 + 22 * m ** 2 + 462 * m + 3 * (42 * m + 2 * m ** 2)


def useful_flops(a16, its, N, L):
    """Per-instance useful flops by the route the library takes (cmpc_api.cu presolve_kind / ipm_kind, automatic
    back-end): presolve = condensed Cholesky up to 20 free leg-steps, stage-wise sweep above; interior point =
    condensed factorisations up to 42 free leg-steps, stage-wise sweeps above.  a16 [B, N*L]: bit 15 = swing leg."""
    stance = ((a16.reshape(len(a16), N, L) & 0x8000) == 0)
    nb = stance.sum(axis=(1, 2))
    n = 3.0 * nb
    sweep = stage_flops(3 * stance.sum(axis=2)).sum(axis=1)          # one backward + forward sweep over the horizon
    pre = np.where(nb <= 20, n ** 3 / 3 + 2 * n ** 2, sweep)         # presolve or final polish system
    it = np.where(nb <= 42, its * (n ** 3 / 3 + 4 * n ** 2), its * sweep)
    return it + pre


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


def cpu_legs(pkg, cfg, st, ds, di, min_seconds=8.0):
    """The reference path on the host cores, same workload, full batch per call: the CPU baseline port (compact
    build, workspace, persistent pool) and, beside it, the explicit oracle (the slow checker)."""
    orc = ge.load_oracle()
    ccfg = pkg.make_config(cfg)
    cores = os.cpu_count() or 1
    B = len(st)
    orc.fast_solve_batch(ccfg, st[:256], ds[:256], di[:256], nthreads=cores)   # warm-up: creates the thread pool
    done, t0 = 0, time.perf_counter()
    while True:
        orc.fast_solve_batch(ccfg, st, ds, di, nthreads=cores)
        done += B
        el = time.perf_counter() - t0
        if el >= min_seconds:
            break
    fast = done / el
    n2 = min(B, 2048)
    t0 = time.perf_counter()
    orc.solve_batch(ccfg, st[:n2], ds[:n2], di[:n2], nthreads=cores, want_lam=False)
    slow = n2 / (time.perf_counter() - t0)
    return {"value": fast, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{done} instances of the same workload ({B}-instance batch x {done // B} calls), {el:.1f} s wall; "
                      "oracle/cmpc_cpu_fast.c: the reference path's CPU port (same presolve + interior point + polish; compact "
                      "closed-form build, no allocation per solve, -O3 -march=native, persistent pthread pool), NOT the "
                      "reference's CasADi/IPOPT binary (not buildable here)",
            "us_per_solve_per_core": 1e6 * cores / fast,
            "explicit_oracle": {"value": slow, "unit": UNIT, "sample": f"{n2} instances, one call; oracle/cmpc_oracle.c, the dense "
                                "explicit checker (thread spawn per call, allocations per solve): NOT a performance baseline"}}


def run_reference(args):
    """--impl reference: the reference path's CPU port, all host threads, the arm's own config and full batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = ge.load_package()
    wl = pkg.workloads
    cfg = dict(wl.default_config(horizon=args.horizon), presolve=args.presolve)
    B = args.batch
    st, ds, di = wl.make_batch(cfg, B, gaits=tuple(args.gaits.split(",")))
    orc = ge.load_oracle()
    ccfg = pkg.make_config(cfg)
    cores = os.cpu_count() or 1
    for _ in range(max(3, args.warmup)):
        orc.fast_solve_batch(ccfg, st, ds, di, nthreads=cores)
    steps = args.steps
    t0 = time.perf_counter()
    for k in range(steps):
        orc.fast_solve_batch(ccfg, st, ds, di, nthreads=cores)
        if time.perf_counter() - t0 > 150.0:   # bounded: a slow host stops early and says so
            steps = k + 1
            break
    el = time.perf_counter() - t0
    value = B * steps / el
    line = {"impl": "reference", "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": max(3, args.warmup), "ms_per_step": 1e3 * el / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "presolve": args.presolve,
                       "note": "one step = the full batch on the host cores (rank 0 only)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{B} instances per step x {steps} steps; oracle/cmpc_cpu_fast.c, the CPU port of the "
                                       "reference path (persistent thread pool), NOT the reference's CasADi/IPOPT binary "
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
    ap.add_argument("--no-extra", action="store_true", help="skip the constrained / config3 / config4 / strong blocks")
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
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def timed_block(bcfg, nb, bgaits, first, steps, warm=3):
        """Device-resident timing of one workload on this rank: median / p99 step time (max over ranks), plus the
        solver statistics of this rank's shard."""
        bst, bds, bdi = wl.make_batch(bcfg, nb, first=first, gaits=bgaits)
        m = pkg.CentroidalMPC.from_dict(bcfg, device=local_rank)
        m.SetupMPC(nb)
        m.set_stream(stream.cuda_stream)
        d = [torch.from_numpy(a).to(dev) for a in (bst, bds, bdi)]
        f = torch.zeros(nb, m.n_forces, dtype=torch.float64, device=dev)
        s = torch.zeros(nb, dtype=torch.int32, device=dev)
        it = torch.zeros(nb, dtype=torch.int32, device=dev)
        kk = torch.zeros(nb, dtype=torch.float64, device=dev)
        act = torch.zeros(nb, m.N * m.L, dtype=torch.int16, device=dev)
        run = lambda: m.solve_device(nb, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), f.data_ptr(), s.data_ptr(),
                                     it.data_ptr(), kk.data_ptr(), 0, act.data_ptr())
        for _ in range(warm):
            run()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for e0, e1 in evs:
            flush.fill_(1)
            e0.record(stream); run(); e1.record(stream)
        barrier()
        ms = np.array([a.elapsed_time(b) for a, b in evs])
        st_ = s.cpu().numpy(); its = it.cpu().numpy(); kkt_ = kk.cpu().numpy()
        a16 = act.cpu().numpy().view(np.uint16)
        NL = (m.N, m.L)
        nact = np.array([[bin(int(x) & 0x3FF).count("1") for x in row] for row in a16[:256]]).sum(axis=1)
        nfree = 3 * ((a16 & 0x8000) == 0).sum(axis=1)
        ok = st_ <= 1
        m.close()
        p50 = reduce_max(float(np.median(ms)))
        return {"p50_ms": p50, "p99_ms": reduce_max(float(np.percentile(ms, 99))), "steps": steps, "batch_per_gpu": nb,
                "status_counts": np.bincount(st_, minlength=5).tolist(), "mean_ipm_iters": float(its.mean()),
                "max_kkt": float(kkt_[ok].max()) if ok.any() else None, "mean_active_rows": float(nact.mean()),
                "flops_per_solve": float(np.mean(useful_flops(a16, its, NL[0], NL[1]))), "_nfree": float(nfree.mean())}

    # rank r owns instance ids [r*B, (r+1)*B): contiguous block split, no inter-GPU traffic
    st, ds, di = wl.make_batch(cfg, B, first=rank * B, gaits=gaits)
    mpc = pkg.CentroidalMPC.from_dict(cfg, device=local_rank)
    mpc.SetupMPC(B)
    mpc.set_stream(stream.cuda_stream)

    d_st, d_ds, d_di = (torch.from_numpy(a).to(dev) for a in (st, ds, di))
    d_forces = torch.zeros(B, mpc.n_forces, dtype=torch.float64, device=dev)
    d_status = torch.zeros(B, dtype=torch.int32, device=dev)
    d_iters = torch.zeros(B, dtype=torch.int32, device=dev)
    d_kkt = torch.zeros(B, dtype=torch.float64, device=dev)

    def step_device():
        mpc.solve_device(B, d_st.data_ptr(), d_ds.data_ptr(), d_di.data_ptr(), d_forces.data_ptr(),
                         d_status.data_ptr(), d_iters.data_ptr(), d_kkt.data_ptr())

    for _ in range(max(3, args.warmup)):
        step_device()
    torch.cuda.synchronize()
    st0 = pkg.CmpcStats()
    mpc.solve_device(B, d_st.data_ptr(), d_ds.data_ptr(), d_di.data_ptr(), d_forces.data_ptr(),
                     d_status.data_ptr(), d_iters.data_ptr(), d_kkt.data_ptr(), stats=st0)
    launches_per_step = int(st0.launches)   # one (pre)solve kernel pair per size class in use
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

    for _ in range(12):        # (the library warms up for two calls, then times its three routes twice each)
        step_e2e()
    e2e_steps = max(10, min(args.steps, 100))
    call_ms = []
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        t1 = time.perf_counter()
        step_e2e()             # synchronous: returns after the last output byte has landed
        call_ms.append(1e3 * (time.perf_counter() - t1))
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_route = mpc.last_route()
    assert np.array_equal(h_forces.numpy(), forces0), "e2e path disagrees with device-resident path"
    sampler.stop_flag = True
    # host link of THIS box, same buffers (what bounds e2e): plain pinned copies with the DMA engines
    link = None
    if rank == 0:
        big = torch.empty_like(d_di)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best_in, best_out = 1e9, 1e9
        for _ in range(5):
            e0.record(stream); big.copy_(h_in[2], non_blocking=True); e1.record(stream); torch.cuda.synchronize()
            best_in = min(best_in, e0.elapsed_time(e1))
            e0.record(stream); h_forces.copy_(d_forces, non_blocking=True); e1.record(stream); torch.cuda.synchronize()
            best_out = min(best_out, e0.elapsed_time(e1))
        link = {"h2d_gbs": h_in[2].numel() * 8 / (best_in * 1e-3) / 1e9, "d2h_gbs": h_forces.numel() * 8 / (best_out * 1e-3) / 1e9}

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

    # ---- the other workloads, each device-resident like `value` (every rank takes part; max over ranks)
    extra = {}
    if not args.no_extra:
        hard10, hard30 = wl.hard_config(10, 0.3), wl.hard_config(30, 0.3)
        r = timed_block(hard10, 4096, wl.GAITS, rank * 4096, 20)
        extra["constrained"] = dict(r, value=world * 4096 / (r["p50_ms"] * 1e-3), unit=UNIT,
                                    workload="tracking-heavy weights (workloads.hard_config), mu 0.3, dt 0.03, horizon 10, mixed gaits, "
                                             "4096 per GPU: friction rows active in every instance; route: presolve defers, "
                                             "interior point + active-set polish (dense classes) ")
        r = timed_block(dict(wl.default_config(30)), 1024, ("trot",), rank * 1024, 20)
        extra["config3"] = dict(r, value=world * 1024 / (r["p50_ms"] * 1e-3), unit=UNIT,
                                workload="BASELINE config 3: horizon 30, batch 1024 per GPU, trot, reference weights (settled by the "
                                         "stage-wise Riccati presolve; the condensed 360 x 360 H is never formed)")
        r = timed_block(hard30, 4096, wl.GAITS, rank * 4096, 5, warm=2)
        extra["config3_constrained"] = dict(r, value=world * 4096 / (r["p50_ms"] * 1e-3), unit=UNIT,
                                            workload="horizon 30, tracking-heavy weights, mu 0.3, mixed gaits, 4096 per GPU: active rows in "
                                                     "every instance; stage-wise (Riccati) interior point + polish kernel")
        per = 65536 // world
        r = timed_block(dict(wl.default_config(10)), per, wl.GAITS, rank * per, 10)
        extra["config4"] = dict(r, value=65536 / (r["p50_ms"] * 1e-3), unit=UNIT,
                                workload=f"BASELINE config 4: 65536 mixed-gait instances in total, {per} per GPU over {world} GPU(s), horizon 10, reference weights")
        if world > 1:
            per = 4096 // world
            r = timed_block(cfg, per, gaits, rank * per, 50)
            extra["strong"] = dict(r, value=4096 / (r["p50_ms"] * 1e-3), unit=UNIT, scaling="strong",
                                   workload=f"ONE headline batch of 4096 split over {world} GPUs ({per} each): a batch this small is "
                                            "tail-dominated (one partial wave per GPU), see config4 for a batch that fills the GPUs")

    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_s_max = float(t[0]), float(t[1])

    if rank == 0:
        n_free = float(np.mean([(np.asarray(di[b]).reshape(4, -1)[:, :args.horizon] > 0).sum() * 3 for b in range(min(B, 256))]))
        mean_it = float(iters.mean())
        flops = float(alg_flops(n_free, mean_it)) * B
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        fp64_peak = mpc.measure_fp64_peak()
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
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
        settled = float((iters == 0).mean())
        line = {
            "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args),
                       "l2": "flushed between timed steps (256 MiB fill)", "sharding": f"dp{world} (independent instances)",
                       "presolve": args.presolve,
                       "solver_path": f"{100 * settled:.1f}% of the instances are settled by the presolve (unconstrained minimiser of the "
                                      "QP verified feasible: one Cholesky of H, no interior-point iteration, no active row); the "
                                      "constrained-QP path is measured in the `constrained` block"},
            "p50_batch_latency_ms": med_ms, "p99_batch_latency_ms": float(np.percentile(step_ms, 99)),
            "mean_ipm_iters": mean_it, "max_kkt": float(kkt.max()),
            "status_counts": np.bincount(status, minlength=5).tolist(),
            "e2e": {"value": world * B * e2e_steps / e2e_s_max, "unit": UNIT,
                    "h2d_bytes_per_step": B * in_bytes, "d2h_bytes_per_step": B * (8 * mpc.n_forces + 4),
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s_max / e2e_steps,
                    "p50_call_ms": float(np.median(call_ms)), "p99_call_ms": float(np.percentile(call_ms, 99)),
                    "route": e2e_route, "host_link_this_box": link,
                    "link_floor_ms": (B * in_bytes / (link["h2d_gbs"] * 1e6)) if link else None},
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
            f2 = float(alg_flops(n_free, ipm_only["mean_ipm_iters"]))
            ipm_only["roofline_frac"] = f2 * B / (ipm_only["p50_batch_latency_ms"] * 1e-3) / 1e12 / fp64_peak
            ipm_only["flops_per_solve"] = f2
            line["ipm_only"] = ipm_only
        for k, r in extra.items():
            nb = r["batch_per_gpu"]
            r["roofline_frac"] = r["flops_per_solve"] * nb / (r["p50_ms"] * 1e-3) / 1e12 / fp64_peak
            r["n_free_mean"] = r.pop("_nfree")
            line[k] = r
        if gather is not None:
            line["nccl_gather"] = gather
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_legs(pkg, cfg, st, ds, di)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    mpc.close()


if __name__ == "__main__":
    main()
