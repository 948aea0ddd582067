"""End-to-end step anatomy: cmpc_solve_batch with pinned buffers, library-side event spans."""
import sys, time, ctypes as C; sys.path.insert(0, '.')
import numpy as np, torch, __graft_entry__ as ge
pkg = ge.load_package(); wl = pkg.workloads
B = 4096
cfg = wl.default_config(10); st, ds, di = wl.make_batch(cfg, B)
m = pkg.CentroidalMPC.from_dict(cfg); m.SetupMPC(B)
pin = [torch.from_numpy(a).pin_memory() for a in (st, ds, di)]
f = torch.zeros(B, m.n_forces, dtype=torch.float64).pin_memory(); s = torch.zeros(B, dtype=torch.int32).pin_memory()
vp = C.c_void_p
def step(stats=None):
    rc = m.lib.cmpc_solve_batch(m.h, B, vp(pin[0].data_ptr()), vp(pin[1].data_ptr()), vp(pin[2].data_ptr()), vp(f.data_ptr()),
                                vp(s.data_ptr()), None, None, None, None, C.byref(stats) if stats else None)
    assert rc == 0
for _ in range(10): step()
t0 = time.perf_counter()
for _ in range(50): step()
wall = (time.perf_counter() - t0) / 50
stt = pkg.CmpcStats(); step(stt)
print(f"wall {wall*1e3:.3f} ms/step; spans: copy-in {stt.h2d_ms:.3f} ms, until last kernel {stt.kernel_ms:.3f} ms, copy-out {stt.d2h_ms:.3f} ms, launches {stt.launches}")
