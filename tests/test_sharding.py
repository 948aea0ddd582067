"""Multi-GPU host logic on CPU: world_size-2 gloo run of the contiguous block split used by
bench.py (rank r owns instance ids [r*B, (r+1)*B)), no data-path collective; the per-rank
results gathered (the optional NCCL gather of SURVEY §8e, here over gloo) equal the
single-process batch bit for bit.  The solver on each rank is the CPU oracle (test-only)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import __graft_entry__ as ge
pkg, orc = ge.load_package(), ge.load_oracle()
wl = pkg.workloads
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
B = 24
cfg = wl.default_config(10)
st, ds, di = wl.make_batch(cfg, B, first=rank * B, gaits=wl.GAITS)
res = orc.solve_batch(pkg.make_config(cfg), st, ds, di)
mine = torch.from_numpy(res["forces"])
parts = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(parts, mine)
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)   # the max-over-ranks timing reduction of bench.py
if rank == 0:
    np.save(sys.argv[2], torch.cat(parts).numpy())
    assert float(t) == world
dist.barrier(); dist.destroy_process_group()
'''


def test_block_split_world2_gloo(tmp_path, pkg, orc, wl):
    out = tmp_path / "gathered.npy"
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), ROOT, str(out)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    gathered = np.load(out)
    cfg = wl.default_config(10)
    st, ds, di = wl.make_batch(cfg, 48, gaits=wl.GAITS)
    whole = orc.solve_batch(pkg.make_config(cfg), st, ds, di)["forces"]
    assert np.array_equal(gathered, whole)


def test_make_batch_is_shard_invariant(wl):
    cfg = wl.default_config(10)
    a = wl.make_batch(cfg, 32, gaits=wl.GAITS)
    b0 = wl.make_batch(cfg, 16, first=0, gaits=wl.GAITS)
    b1 = wl.make_batch(cfg, 16, first=16, gaits=wl.GAITS)
    for x, y0, y1 in zip(a, b0, b1):
        assert np.array_equal(x, np.concatenate([y0, y1]))


def test_gait_tables_always_have_a_stance_leg(wl):
    for g in wl.GAITS:
        for ph in range(10):
            for N in (6, 10, 30):
                t = wl.gait_table(g, N, ph)
                assert t.shape == (4, N) and (t.sum(axis=0) >= 1).all()
