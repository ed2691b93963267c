"""Development probe (the sky-ambient half needs tools/exp/k3_sky_refill_variant.diff applied; without it k3_sky_stream is an unknown option): direct light on the C5 map (sun + 162-direction sky ambient over 2.0 M luxels) with the sky-ambient pass in
its two forms (k3_sky_stream 1 = tiled direction-major refill, 0 = the warp walks the directions in lockstep): device ms and bit-identity of the
result.  Also K1 on S1 with the sort restricted to the key's leading bits (k1_sort_bits).  Writes gpurun_out/r02_k3_sky.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
res = {}
dev = torch.device("cuda", 0)
dirs = np.loadtxt(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vrad_b200", "data", "anorms.txt"), dtype=np.float32)
for name, sc in (("S3_outdoor", scenes.outdoor()),):
    env = environment_from_scene(sc, with_patches=False)
    env.set_sky_dirs(dirs)
    lights = sc.lights
    if not any(int(l["type"]) == 5 for l in lights):
        lights = np.concatenate([lights[:1], lights[:1]]); lights[1]["type"] = 5        # add a sky-ambient light
    d_pos, d_nrm = torch.from_numpy(sc.luxel_pos).to(dev), torch.from_numpy(sc.luxel_normal).to(dev)
    d_rgb = torch.empty((d_pos.shape[0], 3), device=dev)
    ref = None
    for flag in (1, 0):
        env.set_option("k3_sky_stream", flag)
        env.direct_light(d_pos, d_nrm, lights, out=d_rgb)
        env.direct_light(d_pos, d_nrm, lights, out=d_rgb)
        ms, nl = env.last_timing()
        got = d_rgb.cpu().numpy()
        if ref is None: ref = got
        res[f"{name}_sky_stream{flag}"] = {"direct_light_ms": ms, "launches": nl, "luxels": int(d_pos.shape[0]), "identical": bool(np.array_equal(got.view(np.uint32), ref.view(np.uint32)))}
        print(name, "k3_sky_stream", flag, ms, "ms identical", res[f"{name}_sky_stream{flag}"]["identical"], flush=True)
    env.close()
s1 = scenes.box_room(); env = environment_from_scene(s1, with_patches=False)
env.set_stream(torch.cuda.current_stream().cuda_stream); env.set_async(True)
a, b = scenes.shadow_segments(s1, 1 << 24, seed=0xC1)
d_a, d_b = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
bits = torch.empty((1 << 24) // 32, dtype=torch.int32, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ref = None
env.set_option("k1_sort", 1)
for key in (0, 1):
    for sb in (30, 24, 18, 16, 12):
        env.set_option("k1_key", key); env.set_option("k1_sort_bits", sb)
        for _ in range(2): env.test_lines(d_a, d_b, out=bits)
        e0.record()
        for _ in range(5): env.test_lines(d_a, d_b, out=bits)
        e1.record(); torch.cuda.synchronize()
        got = bits.cpu().numpy()
        if ref is None: ref = got
        ms = e0.elapsed_time(e1) / 5
        res[f"k1_s1_key{key}_sortbits{sb}"] = {"ms": ms, "seg_per_s": (1 << 24) / (ms * 1e-3), "same_bits": bool(np.array_equal(got, ref))}
        print("k1 key", key, "sort bits", sb, ms, "ms", (1 << 24) / (ms * 1e-3), flush=True)
env.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r02_k3_sky.json", "w"), indent=1)
