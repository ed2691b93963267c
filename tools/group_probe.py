"""Development probe: the in-process multi-GPU handle (vrad_env_create_multi) on all visible GPUs -- transfer build and bounce
gather of the C4 map, K1 over host buffers.  One process; writes gpurun_out/r02_group_probe_n<N>.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
from vrad_b200.lib import PinnedArray
n_dev = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
devices = list(range(n_dev))
res = {"devices": devices}
s2 = scenes.multi_room(); N = s2.n_patches
env = environment_from_scene(s2, devices=devices)
t0 = time.perf_counter(); nnz = env.build_transfers(s2.pvs); res["transfer_build_s_first"] = time.perf_counter() - t0
t0 = time.perf_counter(); nnz = env.build_transfers(s2.pvs); res["transfer_build_s"] = time.perf_counter() - t0
res["nnz"] = nnz
emit0 = scenes.SplitMix64(0xE1).uniform(3 * N, 0.0, 200.0).reshape(N, 3)
out = np.empty_like(emit0)
env.bounce(emit0, 100, out=out, want_added=False)
t0 = time.perf_counter()
for _ in range(5):
    env.bounce(emit0, 100, out=out, want_added=False)
dt = (time.perf_counter() - t0) / 5
ms, launches = env.last_timing()
res["bounce_100_wall_ms"] = dt * 1e3; res["bounce_100_device_ms_max_rank"] = ms; res["iters_per_sec_wall"] = 100 / dt
single = environment_from_scene(s2, device=0); single.build_transfers(s2.pvs)
ref, _, _ = single.bounce(emit0, 100); single.close()
res["max_rel_vs_single_gpu"] = float(np.abs(out - ref).max() / np.abs(ref).max())
env.close()
s1 = scenes.box_room()
env = environment_from_scene(s1, devices=devices, with_patches=False)
n = 1 << 24
pts, pairs = scenes.shadow_segment_indices(s1, n, seed=0xC0FFEE)
env.points_upload(pts)
hp = PinnedArray((n, 2), np.int32); hp.array[...] = pairs; hb = PinnedArray((n // 32,), np.uint32)
env.test_lines_indexed(hp.array, out=hb.array)
t0 = time.perf_counter()
for _ in range(3):
    env.test_lines_indexed(hp.array, out=hb.array)
res["rays_indexed_host_per_sec"] = 3 * n / (time.perf_counter() - t0)
one = environment_from_scene(s1, device=0, with_patches=False); one.points_upload(pts)
res["rays_bits_equal_single_gpu"] = bool(np.array_equal(hb.array, one.test_lines_indexed(pairs))); one.close()
env.close()
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/r02_group_probe_n{n_dev}.json", "w"), indent=1)
