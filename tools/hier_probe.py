"""Development probe: the full S2 map with its patch hierarchy -- transfer count, build time, bounce time.
VRAD_K4_SHORT / VRAD_K4_SHORT_CFG select the gather form (read once per process)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene

t0 = time.time(); sc = scenes.multi_room_hier(nx=12, ny=11); t = sc.meta["tree"]
leaf = t["child1"] == -1
print(f"scene {time.time()-t0:.1f}s: patches {sc.n_patches} leaves {leaf.sum()} roots {(t['parent']==-1).sum()}")
env = environment_from_scene(sc)
env.set_hierarchy(t["parent"], t["child1"], t["child2"], t["face"])
env.build_transfers(sc.pvs)
t0 = time.time(); nnz = env.build_transfers(sc.pvs); wall = time.time() - t0
print(f"hier build_transfers: nnz={nnz} wall {wall:.3f}s kernels {env.last_timing()}")
N = sc.n_patches
emit0 = torch.full((N, 3), 100.0, device="cuda"); tot = torch.empty_like(emit0)
env.set_stream(torch.cuda.current_stream().cuda_stream); env.set_async(True)
nb = 100
ts = []
for _ in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); env.bounce(emit0, nb, out=tot, want_added=False); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
best = min(ts[1:])
by = 8 * nnz + 40 * N
print(f"hier bounce [{os.environ.get('VRAD_K4_SHORT','auto')}/{os.environ.get('VRAD_K4_SHORT_CFG','88')}]: {best/nb*1e3:.1f} us/bounce "
      f"({nb/(best/1e3):.0f} iters/s), {by*nb/(best/1e3)/1e9:.0f} GB/s algorithmic, launches {env.last_timing()[1]}")
print("total checksum", float(tot.double().sum()), tot.mean(0).tolist())
