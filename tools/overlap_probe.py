"""Why does the route overlap (CMPC_OVERLAP) show under bench.py's timing loop and not under a per-step synchronised
one?  Same workload, four timing styles."""
import sys, json; sys.path.insert(0, '.')
import numpy as np, torch, __graft_entry__ as ge
pkg = ge.load_package(); wl = pkg.workloads
cfg = wl.hard_config(10, 0.3); B = 4096
st, ds, di = wl.make_batch(cfg, B, gaits=wl.GAITS)
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def run_case(sync_each, use_flush, own_stream, act_out):
    stream = torch.cuda.Stream()
    m = pkg.CentroidalMPC.from_dict(cfg); m.SetupMPC(B)
    if not own_stream: m.set_stream(stream.cuda_stream)
    d = [torch.from_numpy(a).to(dev) for a in (st, ds, di)]
    f = torch.zeros(B, m.n_forces, dtype=torch.float64, device=dev); s = torch.zeros(B, dtype=torch.int32, device=dev)
    it = torch.zeros(B, dtype=torch.int32, device=dev); kk = torch.zeros(B, dtype=torch.float64, device=dev)
    act = torch.zeros(B, m.N * m.L, dtype=torch.int16, device=dev)
    torch.cuda.synchronize()
    run = lambda: m.solve_device(B, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), f.data_ptr(), s.data_ptr(), it.data_ptr(), kk.data_ptr(),
                                 0, act.data_ptr() if act_out else 0)
    for _ in range(3): run()
    ms = []
    with torch.cuda.stream(stream):
        if own_stream:
            import time
            for _ in range(10):
                m.synchronize(); t0 = time.perf_counter(); run(); m.synchronize(); ms.append((time.perf_counter() - t0) * 1e3)
        else:
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
            import time; th = time.perf_counter()
            for e0, e1 in evs:
                if use_flush: flush.fill_(1)
                e0.record(stream); run(); e1.record(stream)
                if sync_each: torch.cuda.synchronize()
            print(f"   host loop {1e3 * (time.perf_counter() - th):.2f} ms", end="")
            torch.cuda.synchronize()
            ms = [a.elapsed_time(b) for a, b in evs]
    m.close()
    return float(np.median(ms)), [round(float(x), 2) for x in ms]

import os
cases = (("bench-like (flush, no sync)", (False, True, False, True)), ("no flush, no sync", (False, False, False, True)),
                   ("flush, sync each", (True, True, False, True)), ("no flush, sync each", (True, False, False, True)),
                   ("no act output, sync each", (True, False, False, False)), ("own stream, wall clock", (True, False, True, True)))
for name, args in cases[:int(os.environ.get("NCASES", "6"))]:
    med, all_ms = run_case(*args)
    print(f"{name:32s} {med:.3f} ms  {all_ms}", flush=True)
