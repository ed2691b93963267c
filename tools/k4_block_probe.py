"""Development probe: the C4 gather from the block-row streams (k4_pack 2: 4 rows share a column list) against the packed streams (1)
and the pairs (0) on one GPU.  (The committed profiles/r02_k4_block.json also holds the kernel forms tried while the kernel was written --
entries in flight x resident blocks per SM, "variant" 1..6 -- the shipped form is variant 4.)  gpurun_out/r02_k4_block.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
dev = torch.device("cuda", 0)
res = {}
s = scenes.multi_room()
env = environment_from_scene(s); env.set_stream(torch.cuda.current_stream().cuda_stream)
nnz = env.build_transfers(s.pvs)
N = s.n_patches
e0 = torch.from_numpy(scenes.SplitMix64(0xE1).uniform(3 * N, 0.0, 200.0).reshape(N, 3)).to(dev); out = torch.empty_like(e0)
env.set_async(True)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ref = None
for pack, var in ((0, 0), (1, 0), (2, 0), (2, 0), (1, 0), (0, 0)):
    t0 = time.perf_counter(); env.set_option("k4_pack", pack); torch.cuda.synchronize(); replan = time.perf_counter() - t0
    t1, a1, _ = env.bounce(e0, 5)
    t1, a1 = (np.asarray(v.cpu()) if hasattr(v, 'cpu') else np.asarray(v) for v in (t1, a1))
    env.bounce(e0, 100, out=out, want_added=False)
    torch.cuda.synchronize(); ev0.record()
    for _ in range(3): env.bounce(e0, 100, out=out, want_added=False)
    ev1.record(); torch.cuda.synchronize()
    us = ev0.elapsed_time(ev1) / 300 * 1e3
    got = out.cpu().numpy()
    if ref is None: ref = (got, t1, a1)
    key = f"pack{pack}_variant{var}"
    res[key + ("_again" if key in res else "")] = {"us_per_bounce": us, "replan_s": replan, "layout": env.transfers_layout(), "max_rel_vs_pairs": float(np.abs(got - ref[0]).max() / np.abs(ref[0]).max()),
                                                  "five_bounces_max_rel": float(np.abs(t1 - ref[1]).max() / np.abs(ref[1]).max()), "added_rel": float(np.abs(a1 - ref[2]).max() / np.abs(ref[2]).max())}
    print("k4_pack", pack, "variant", var, round(us, 2), "us/bounce", env.transfers_layout(), res[key + ("_again" if key + "_again" in res else "")]["max_rel_vs_pairs"],
          res[key + ("_again" if key + "_again" in res else "")]["added_rel"], "replan", round(replan, 3), flush=True)
env.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r02_k4_block.json", "w"), indent=1)
