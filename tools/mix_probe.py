import sys, json; sys.path.insert(0,'.'); sys.path.insert(0,'tools')
import ripm_probe as rp, __graft_entry__ as ge
pkg=ge.load_package(); wl=pkg.workloads
cfg=wl.hard_config(10,0.3)
for gaits in (wl.GAITS, ("stand",)):
    st,ds,di=wl.make_batch(cfg,4096,gaits=gaits)
    print(gaits[:2], json.dumps(rp.timed(pkg,cfg,st,ds,di,steps=10)))
