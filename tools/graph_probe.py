"""Capture the device-resident solve (all kernels of a call, fork / join included) into a CUDA graph and replay it:
results must equal the eager call bit for bit; prints eager vs replay time per step."""
import sys; sys.path.insert(0, '.')
import numpy as np, torch, __graft_entry__ as ge
pkg = ge.load_package(); wl = pkg.workloads
dev = torch.device("cuda", 0)
for name, cfg, gaits in (("headline", wl.default_config(10), ("trot",)), ("constrained mixed", wl.hard_config(10, 0.3), wl.GAITS)):
    B = 4096
    st, ds, di = wl.make_batch(cfg, B, gaits=gaits)
    m = pkg.CentroidalMPC.from_dict(cfg); m.SetupMPC(B)
    stream = torch.cuda.Stream()
    m.set_stream(stream.cuda_stream)
    d = [torch.from_numpy(a).to(dev) for a in (st, ds, di)]
    f = torch.zeros(B, m.n_forces, dtype=torch.float64, device=dev); s = torch.zeros(B, dtype=torch.int32, device=dev)
    it = torch.zeros(B, dtype=torch.int32, device=dev); kk = torch.zeros(B, dtype=torch.float64, device=dev)
    run = lambda: m.solve_device(B, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), f.data_ptr(), s.data_ptr(), it.data_ptr(), kk.data_ptr())
    for _ in range(3):
        run(); torch.cuda.synchronize()
    f_eager = f.clone(); it_eager = it.clone()
    def timed(fn, n=20):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        with torch.cuda.stream(stream):
            e[0].record(stream)
            for i in range(n):
                fn(); e[i + 1].record(stream)
        torch.cuda.synchronize()
        return float(np.median([e[i].elapsed_time(e[i + 1]) for i in range(n)]))
    t_eager = timed(run)
    g = torch.cuda.CUDAGraph()
    f.zero_(); it.zero_()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=stream):
        run()
    g.replay(); torch.cuda.synchronize()
    same = bool(torch.equal(f, f_eager) and torch.equal(it, it_eager))
    t_graph = timed(g.replay)
    print(f"{name}: eager {t_eager:.4f} ms, graph replay {t_graph:.4f} ms, identical results: {same}", flush=True)
    m.close()
