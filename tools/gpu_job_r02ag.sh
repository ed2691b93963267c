set -x
python - <<'PY'
import sys, json
sys.path.insert(0, '.')
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
dev = torch.device("cuda", 0)
s = scenes.multi_room(); N = s.n_patches
e0 = torch.from_numpy(scenes.SplitMix64(0xE1).uniform(3 * N, 0.0, 200.0).reshape(N, 3)).to(dev); out = torch.empty_like(e0)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {}
for world, rank in ((8, 3), (8, 6), (4, 1)):
    env = environment_from_scene(s, rank=rank, world=world); env.set_stream(torch.cuda.current_stream().cuda_stream)
    env.set_option("k4_sim_peers", 1)
    nnz = env.build_transfers(s.pvs); env.set_async(True)
    for pack, bps, pool in ((3, 0, 25), (3, 1, 25), (3, 2, 25), (3, 3, 25), (3, 4, 25), (3, 2, 12), (3, 2, 40), (3, 3, 12), (1, 0, 25), (3, 0, 25), (1, 0, 25)):
        env.set_option("k4_pool", pool); env.set_option("k4_bk_bps", bps); env.set_option("k4_pack", pack)
        env.bounce(e0, 100, out=out, want_added=False)
        torch.cuda.synchronize(); ev0.record()
        for _ in range(3): env.bounce(e0, 100, out=out, want_added=False)
        ev1.record(); torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) / 300 * 1e3
        key = f"world{world}_rank{rank}_pack{pack}_bps{bps}_pool{pool}"
        res[key + ("_again" if key in res else "")] = us
        print("world", world, "rank", rank, "pack", pack, "bps", bps, "pool", pool, round(us, 2), "us/bounce", flush=True)
    env.close()
json.dump(res, open("gpurun_out/r02_k4_block_bps_sim.json", "w"), indent=1)
PY
