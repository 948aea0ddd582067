import sys; sys.path.insert(0, '.')
import numpy as np, __graft_entry__ as ge
pkg = ge.load_package(); wl = pkg.workloads
for N, gaits in ((30, ("trot",)), (30, ("trot", "stand", "gallop")), (20, ("trot",)), (16, ("stand",))):
    cfg = wl.default_config(N); B = 256
    st, ds, di = wl.make_batch(cfg, B, gaits=gaits)
    m = pkg.CentroidalMPC.from_dict(cfg); m.SetupMPC(B)
    out = m.UpdateMPCBatch(st, ds, di)
    nb = (di.reshape(B, 4, -1)[:, :, :N] > 0).sum(axis=(1, 2))
    it = out["iters"]
    print(N, gaits, "deferred", (it > 0).sum(), "of", B, "nb of deferred", np.unique(nb[it > 0]), "nb all", np.unique(nb), "kkt", out["kkt"].max())
    m.close()
