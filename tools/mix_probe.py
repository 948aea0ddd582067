"""Constrained workload (hard_config), device-resident: throughput of each gait alone and of the mix.
   python tools/mix_probe.py [horizon]"""
import sys, json; sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import ripm_probe as rp, __graft_entry__ as ge
pkg = ge.load_package(); wl = pkg.workloads
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
cfg = wl.hard_config(N, 0.3)
only_mix = len(sys.argv) > 2 and sys.argv[2] == "mix"
for gaits in ([] if only_mix else [(g,) for g in wl.GAITS]) + [wl.GAITS]:
    st, ds, di = wl.make_batch(cfg, 4096, gaits=gaits)
    print("+".join(gaits), json.dumps(rp.timed(pkg, cfg, st, ds, di, steps=10)), flush=True)
    if gaits == wl.GAITS:
        print("  (legacy default stream)", json.dumps(rp.timed(pkg, cfg, st, ds, di, steps=10, default_stream=True)), flush=True)
