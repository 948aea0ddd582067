"""End-to-end anatomy per route: wall time per call and the library's event timeline (CMPC_DEBUG_TIMELINE) of one call.
   python tools/e2e_timeline.py [mode ...]      (CMPC_E2E_MODE values; CMPC_E2E_CHUNK from the environment)"""
import os, sys, time, ctypes as C; sys.path.insert(0, '.')
modes = [int(a) for a in sys.argv[1:]] or [1, 4, 5, 3]
import numpy as np, torch, __graft_entry__ as ge
pkg = ge.load_package(); wl = pkg.workloads
B = 4096
cfg = wl.default_config(10); st, ds, di = wl.make_batch(cfg, B)
pin = [torch.from_numpy(a).pin_memory() for a in (st, ds, di)]
vp = C.c_void_p
for mode in modes:
    os.environ["CMPC_E2E_MODE"] = str(mode)
    os.environ.pop("CMPC_DEBUG_TIMELINE", None)
    m = pkg.CentroidalMPC.from_dict(cfg); m.SetupMPC(B)
    f = torch.zeros(B, m.n_forces, dtype=torch.float64).pin_memory(); s = torch.zeros(B, dtype=torch.int32).pin_memory()
    def step():
        rc = m.lib.cmpc_solve_batch(m.h, B, vp(pin[0].data_ptr()), vp(pin[1].data_ptr()), vp(pin[2].data_ptr()), vp(f.data_ptr()),
                                    vp(s.data_ptr()), None, None, None, None, None)
        assert rc == 0
    for _ in range(20): step()
    ts = []
    for _ in range(100):
        t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e3
    print(f"mode {mode} chunk {os.environ.get('CMPC_E2E_CHUNK', '-')}: wall p50 {np.median(ts):.3f} ms p10 {np.percentile(ts, 10):.3f} p90 {np.percentile(ts, 90):.3f}  route {m.last_route()}", flush=True)
    m.close()
    os.environ["CMPC_DEBUG_TIMELINE"] = "1"
    m = pkg.CentroidalMPC.from_dict(cfg); m.SetupMPC(B)
    for _ in range(4): step()
    m.close()
