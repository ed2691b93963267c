set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_round2.py tests/test_gpu_radiosity.py tests/test_gpu_pipeline.py tests/test_gpu_group.py tests/test_gpu_bump.py tests/test_gpu_fullsize.py tests/test_gpu_cpp_driver.py -x -q 2>&1 | tail -8
python - <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
dev = torch.device("cuda", 0)
s = scenes.multi_room(); N = s.n_patches
e0 = torch.from_numpy(scenes.SplitMix64(0xE1).uniform(3 * N, 0.0, 200.0).reshape(N, 3)).to(dev); out = torch.empty_like(e0)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ref = None
# N=1 plain / work-item kernel on one device, and the rank-3 slice of a simulated 8-rank run
for world, rank, items, pack in ((1, 0, 0, 2), (1, 0, 1, 2), (1, 0, 0, 1), (8, 3, 0, 2), (8, 3, 0, 1), (8, 3, 0, 0)):
    env = environment_from_scene(s, rank=rank, world=world); env.set_stream(torch.cuda.current_stream().cuda_stream)
    if world > 1: env.set_option("k4_sim_peers", 1)
    env.set_option("k4_items", items); env.set_option("k4_pack", pack)
    nnz = env.build_transfers(s.pvs); env.set_async(True)
    env.bounce(e0, 100, out=out, want_added=False)
    torch.cuda.synchronize(); ev0.record()
    for _ in range(3): env.bounce(e0, 100, out=out, want_added=False)
    ev1.record(); torch.cuda.synchronize()
    got = out.cpu().numpy()
    if ref is None: ref = got
    print("world", world, "items", items, "pack", pack, round(ev0.elapsed_time(ev1) / 300 * 1e3, 2), "us/bounce", env.transfers_info(), env.transfers_layout(),
          "max_rel", float(np.abs(got - ref).max() / np.abs(ref).max()) if world == 1 else None, flush=True)
    env.close()
PY
