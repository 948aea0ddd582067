"""Host-link probes (pinned host <-> device, CUDA events): the floor of the end-to-end step.
1. bandwidth and fixed cost per copy; 2. do the fixed costs of copies on different streams overlap?"""
import torch
dev = torch.device("cuda", 0)
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for mb in (0.17, 1.2, 9.57, 64.0):
    n = int(mb * 1e6) // 8
    h = torch.empty(n, dtype=torch.float64).pin_memory(); d = torch.empty(n, dtype=torch.float64, device=dev)
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        us = timed(fn)
        print(f"{name} {mb:6.2f} MB: {us:8.1f} us  {mb*1e3/us:6.2f} GB/s")
# the three input arrays of a 4096-instance batch
sizes = [4096 * 21, 4096 * 99, 4096 * 172]
hs = [torch.empty(n, dtype=torch.float64).pin_memory() for n in sizes]
ds = [torch.empty(n, dtype=torch.float64, device=dev) for n in sizes]
streams = [torch.cuda.Stream() for _ in range(3)]
main = torch.cuda.current_stream()
def run(nstreams, nchunks):
    def fn():
        for k in range(3):
            st = streams[k % nstreams]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                per = (sizes[k] + nchunks - 1) // nchunks
                for c in range(nchunks):
                    ds[k][c * per:(c + 1) * per].copy_(hs[k][c * per:(c + 1) * per], non_blocking=True)
        for st in streams[:nstreams]:
            main.wait_stream(st)
    return timed(fn)
for ns in (1, 3):
    for nc in (1, 2, 4):
        print(f"inputs of one batch (9.57 MB in 3 arrays): {ns} stream(s) x {nc} chunk(s): {run(ns, nc):7.1f} us")
