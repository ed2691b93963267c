"""Development probe: K3 direct light on the full S3 map (2.0 M luxels, sun + 162-direction sky ambient)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
s = scenes.outdoor(); env = environment_from_scene(s, with_patches=False)
dirs = np.loadtxt(os.path.join(os.path.dirname(__file__), "..", "vrad_b200", "data", "anorms.txt"), dtype=np.float32)
env.set_sky_dirs(dirs)
pos, nrm = torch.from_numpy(s.luxel_pos).cuda(), torch.from_numpy(s.luxel_normal).cuda()
out = torch.empty((pos.shape[0], 3), device="cuda")
for lights, name in ((s.lights[:1], "sun only"), (s.lights, "sun + sky ambient")):
    env.direct_light(pos, nrm, lights, out=out); torch.cuda.synchronize()
    t = time.perf_counter(); env.direct_light(pos, nrm, lights, out=out); torch.cuda.synchronize(); dt = time.perf_counter() - t
    ms, nl = env.last_timing()
    up = (dirs @ s.luxel_normal[::97].T > 0.001).sum() / s.luxel_normal[::97].shape[0]
    rays = pos.shape[0] * (1 + (up if len(lights) > 1 else 0))
    print(f"{name}: wall {dt*1e3:.1f} ms, kernels {ms:.1f} ms ({nl} launches), ~{rays/1e6:.0f} M rays -> {rays/(ms*1e-3)/1e9:.2f} G rays/s; lit fraction {(out.sum(1) > 0).float().mean().item():.3f}")
