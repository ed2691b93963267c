"""Development probe: the C5 gather on ONE GPU (5.72e9 transfers, rows over ~3 column windows) from the packed streams (k4_pack 9 = pack
whatever the segment count) against the pairs.  gpurun_out/r02_c5_pack_n1.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
dev = torch.device("cuda", 0)
s3 = scenes.outdoor(); N = s3.n_patches
env = environment_from_scene(s3); env.set_stream(torch.cuda.current_stream().cuda_stream)
t0 = time.perf_counter(); nnz = env.build_transfers(s3.pvs); print("build", time.perf_counter() - t0, nnz, flush=True)
e0 = torch.full((N, 3), 100.0, device=dev); out = torch.empty_like(e0)
env.set_async(True)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {"nnz": int(nnz)}
ref = None
for pack in (0, 9, 0, 9):
    t0 = time.perf_counter(); env.set_option("k4_pack", pack); torch.cuda.synchronize(); replan = time.perf_counter() - t0
    env.bounce(e0, 2, out=out, want_added=False)
    torch.cuda.synchronize(); ev0.record()
    env.bounce(e0, 10, out=out, want_added=False)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 10
    got = out.cpu().numpy()
    if ref is None: ref = got
    lay = env.transfers_layout()
    res[f"pack{pack}" + ("_again" if f"pack{pack}" in res else "")] = {"ms_per_bounce": ms, "replan_s": replan, "layout": lay, "max_rel_vs_pairs": float(np.abs(got - ref).max() / np.abs(ref).max())}
    print("k4_pack", pack, ms, "ms/bounce", lay, float(np.abs(got - ref).max() / np.abs(ref).max()), "replan", replan, flush=True)
env.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r02_c5_pack_n1.json", "w"), indent=1)
