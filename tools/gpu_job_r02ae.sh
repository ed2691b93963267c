set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -k "block or packed or items" 2>&1 | tail -3
VRAD_K4_BK_ROWS=2 timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -k "block_row" 2>&1 | tail -3
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
for world, rank, items in ((8, 3, 0), (4, 1, 0), (2, 1, 0), (1, 0, 1), (1, 0, 0)):
    env = environment_from_scene(s, rank=rank, world=world); env.set_stream(torch.cuda.current_stream().cuda_stream)
    if world > 1: env.set_option("k4_sim_peers", 1)
    env.set_option("k4_items", items)
    nnz = env.build_transfers(s.pvs); env.set_async(True)
    for pack, rows, pool in ((3, 4, 25), (3, 2, 25), (3, 2, 12), (3, 2, 40), (1, 4, 25)):
        env.set_option("k4_pool", pool); env.set_option("k4_bk_rows", rows); env.set_option("k4_pack", pack)
        env.bounce(e0, 100, out=out, want_added=False)
        torch.cuda.synchronize(); ev0.record()
        for _ in range(3): env.bounce(e0, 100, out=out, want_added=False)
        ev1.record(); torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) / 300 * 1e3
        res[f"world{world}_items{items}_pack{pack}_rows{rows}_pool{pool}"] = us
        print("world", world, "items", items, "pack", pack, "rows", rows, "pool", pool, round(us, 2), "us/bounce", env.transfers_layout(), flush=True)
    env.close()
json.dump(res, open("gpurun_out/r02_k4_block_rows_sim.json", "w"), indent=1)
PY
