"""Development probe: the C5 bake on the full S3 map -- K3 direct light, K2 transfers (tile PVS), K4 bounces."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
t = time.time(); s = scenes.outdoor(); print(f"scene {time.time()-t:.1f}s tris {s.n_tris} patches {s.n_patches} clusters {s.n_clusters}", flush=True)
t = time.time(); env = environment_from_scene(s); print(f"build+upload {time.time()-t:.1f}s", flush=True)
free0 = torch.cuda.mem_get_info()[0]
t = time.time(); nnz = env.build_transfers(s.pvs); torch.cuda.synchronize(); dt = time.time() - t
ms, nl = env.last_timing()
print(f"K2: nnz {nnz} ({nnz / s.n_patches:.0f} per row) wall {dt:.2f}s kernels {ms:.0f} ms, device memory used {(free0 - torch.cuda.mem_get_info()[0]) / 1e9:.1f} GB", flush=True)
N = s.n_patches
e0 = torch.full((N, 3), 100.0, device="cuda"); out = torch.empty_like(e0)
env.bounce(e0, 2, out=out, want_added=False); torch.cuda.synchronize()
t = time.time(); env.bounce(e0, 10, out=out, want_added=False); torch.cuda.synchronize(); dt = time.time() - t
print(f"K4: 10 bounces {dt*1e3:.1f} ms -> {dt*100:.2f} ms/bounce, {(8*nnz+40*N)*10/dt/1e9:.0f} GB/s; total finite {bool(torch.isfinite(out).all())} max {out.max().item():.2f}", flush=True)
tot, added, done = env.bounce(e0, 100, early_out=True, out=out)
print(f"early-out bounces {done}, added {added}", flush=True)
