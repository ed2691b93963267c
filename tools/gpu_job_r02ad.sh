set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q 2>&1 | tail -4
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
for world, rank in ((8, 3), (4, 1), (2, 1), (1, 0)):
    env = environment_from_scene(s, rank=rank, world=world); env.set_stream(torch.cuda.current_stream().cuda_stream)
    if world > 1: env.set_option("k4_sim_peers", 1)
    else: env.set_option("k4_items", 1)
    nnz = env.build_transfers(s.pvs); env.set_async(True)
    for pack, pool in ((2, 25), (3, 25), (4, 25), (4, 12), (4, 40), (1, 25)):
        env.set_option("k4_pool", pool); env.set_option("k4_pack", pack)
        env.bounce(e0, 100, out=out, want_added=False)
        torch.cuda.synchronize(); ev0.record()
        for _ in range(3): env.bounce(e0, 100, out=out, want_added=False)
        ev1.record(); torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) / 300 * 1e3
        res[f"world{world}_pack{pack}_pool{pool}"] = us
        print("world", world, "pack", pack, "pool", pool, round(us, 2), "us/bounce", env.transfers_layout(), flush=True)
    env.close()
json.dump(res, open("gpurun_out/r02_k4_block_sim.json", "w"), indent=1)
PY
