"""Development probe: the C4 gather from the packed 6-byte streams (k4_pack 1) against the {col,w} pairs (0) on one GPU: us per
bounce (CUDA events, 3 x 100 bounces), agreement of the two results, and the packed streams' size.  gpurun_out/r02_k4_pack.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
dev = torch.device("cuda", 0)
res = {}
for name, s in (("C4", scenes.multi_room()),):
    env = environment_from_scene(s); env.set_stream(torch.cuda.current_stream().cuda_stream)
    nnz = env.build_transfers(s.pvs)
    N = s.n_patches
    e0 = torch.from_numpy(scenes.SplitMix64(0xE1).uniform(3 * N, 0.0, 200.0).reshape(N, 3)).to(dev); out = torch.empty_like(e0)
    env.set_async(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ref = None
    for pack in (1, 0, 1, 0):
        env.set_option("k4_pack", pack)
        env.bounce(e0, 100, out=out, want_added=False)
        torch.cuda.synchronize(); ev0.record()
        for _ in range(3): env.bounce(e0, 100, out=out, want_added=False)
        ev1.record(); torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) / 300 * 1e3
        got = out.cpu().numpy()
        if ref is None: ref = got
        res[f"{name}_pack{pack}" + ("_again" if f"{name}_pack{pack}" in res else "")] = {"us_per_bounce": us, "nnz": int(nnz), "max_rel_vs_packed": float(np.abs(got - ref).max() / np.abs(ref).max())}
        print(name, "k4_pack", pack, us, "us/bounce", float(np.abs(got - ref).max() / np.abs(ref).max()), flush=True)
    env.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r02_k4_pack.json", "w"), indent=1)
