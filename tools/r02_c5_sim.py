"""C5 (S3 outdoor map) gather: the rank-3 slice of a simulated 8-rank run on ONE GPU (k4_sim_peers) for the work-item kernel's
item size, against the streaming time of the slice.  Writes gpurun_out/r02_c5_sim.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
dev = torch.device("cuda", 0)
s3 = scenes.outdoor(); N = s3.n_patches
env = environment_from_scene(s3, rank=3, world=8)
env.set_stream(torch.cuda.current_stream().cuda_stream)
env.set_option("k4_sim_peers", 1)
t0 = time.perf_counter(); nnz = env.build_transfers(s3.pvs); k2 = time.perf_counter() - t0
row0, row1, _ = env.transfers_info()
res = {"nnz_local": nnz, "rows": [row0, row1], "transfer_build_s": k2, "stream_us_at_peak": 8 * nnz / 6455.3e9 * 1e6}
lens = []
for r0 in range(row0, row1, max(1, (row1 - row0) // 8)):
    rp, _, _ = env.transfers_download_rows(r0, min(row1, r0 + 1024), 1 << 24)
    lens.append(np.diff(rp))
lens = np.concatenate(lens)
res["row_length_sample"] = {"mean": float(lens.mean()), "p50": float(np.percentile(lens, 50)), "p99": float(np.percentile(lens, 99)), "max": int(lens.max()),
                            "frac_gt_2048": float((lens > 2048).mean()), "frac_gt_8192": float((lens > 8192).mean())}
print(res, flush=True)
e0 = torch.full((N, 3), 100.0, device=dev); out = torch.empty_like(e0)
env.set_async(True)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for seg, pool, blk, pack in ((16384, 12, 192, 0), (16384, 12, 192, 1), (16384, 25, 192, 1), (16384, 25, 192, 0), (4096, 12, 192, 1), (16384, 12, 256, 1)):
    env.set_option("k4_seg", seg); env.set_option("k4_pool", pool); env.set_option("k4_block", blk); env.set_option("k4_pack", pack)
    env.bounce(e0, 4, out=out, want_added=False)
    torch.cuda.synchronize(); ev0.record()
    env.bounce(e0, 20, out=out, want_added=False)
    ev1.record(); torch.cuda.synchronize()
    us = ev0.elapsed_time(ev1) / 20 * 1e3
    res[f"seg{seg}_pool{pool}_block{blk}_pack{pack}"] = {"us_per_bounce": us, "frac_of_hbm_peak": (8 * nnz + 40 * (row1 - row0) + 12 * N) / us / 1e3 / 6455.3}
    print(seg, pool, blk, pack, us, float(out.sum().item()), flush=True)
env.close()
json.dump(res, open("gpurun_out/r02_c5_sim_pack.json", "w"), indent=1)
print(json.dumps(res))
