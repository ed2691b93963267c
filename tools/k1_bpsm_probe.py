"""Development probe: resident blocks per SM of the streaming traversal kernel (k1_bpsm 12 = 40 registers with 24 B of spills, 10 = 46,
8 = 55) on S1 (ordered and unordered, 2^24 C1 segments) and S3 (2^22 random segments).  gpurun_out/r02_k1_bpsm.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
dev = torch.device("cuda", 0)
res = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, sc, n, seed in (("S1", scenes.box_room(), 1 << 24, 0xC1), ("S3", scenes.outdoor(), 1 << 22, 0xC5)):
    env = environment_from_scene(sc, with_patches=False)
    env.set_stream(torch.cuda.current_stream().cuda_stream); env.set_async(True)
    a, b = scenes.shadow_segments(sc, n, seed=seed)
    d_a, d_b = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    bits = torch.empty(n // 32, dtype=torch.int32, device=dev)
    ref = None
    for sort in (1, 0):
        for bpsm in (12, 10, 8, 12, 10):
            env.set_option("k1_sort", sort); env.set_option("k1_bpsm", bpsm)
            for _ in range(2): env.test_lines(d_a, d_b, out=bits)
            e0.record()
            for _ in range(5): env.test_lines(d_a, d_b, out=bits)
            e1.record(); torch.cuda.synchronize()
            got = bits.cpu().numpy()
            if ref is None: ref = got
            ms = e0.elapsed_time(e1) / 5
            key = f"{name}_sort{sort}_bpsm{bpsm}"
            res[key + ("_again" if key in res else "")] = {"ms": ms, "seg_per_s": n / (ms * 1e-3), "same_bits": bool(np.array_equal(got, ref))}
            print(name, "sort", sort, "bpsm", bpsm, round(ms, 4), "ms", round(n / (ms * 1e-3) / 1e9, 3), "e9/s", bool(np.array_equal(got, ref)), flush=True)
    env.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r02_k1_bpsm.json", "w"), indent=1)
