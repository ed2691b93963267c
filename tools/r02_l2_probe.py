"""L2 residency of the transfer stream (k4_l2_mb) on the simulated 8-/4-rank slices of the C4 matrix, one GPU."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["VRAD_VERBOSE"] = "1"
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
dev = torch.device("cuda", 0)
s2 = scenes.multi_room(); N = s2.n_patches
e0 = torch.full((N, 3), 100.0, device=dev); out = torch.empty_like(e0)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {}
for world, rank in ((8, 3), (4, 1), (2, 1)):
    env = environment_from_scene(s2, rank=rank, world=world)
    env.set_stream(torch.cuda.current_stream().cuda_stream); env.set_option("k4_sim_peers", 1)
    nnz = env.build_transfers(s2.pvs); env.set_async(True)
    for mb in (0, 32, 64, 96, 128, -64, -96, 0):
        env.set_option("k4_l2_mb", mb)
        env.bounce(e0, 100, out=out, want_added=False); torch.cuda.synchronize()
        ev0.record()
        for _ in range(3): env.bounce(e0, 100, out=out, want_added=False)
        ev1.record(); torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) / 300 * 1e3
        res[f"world{world}_l2mb{mb}"] = us
        print(world, mb, us, 8 * nnz / 1e6, flush=True)
    env.close()
json.dump(res, open("gpurun_out/r02_l2_probe.json", "w"), indent=1)
