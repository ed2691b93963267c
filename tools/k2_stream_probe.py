"""Development probe: the C4 transfer build with pass A in its two forms (k2_stream 1 = compacted ray queues with lane refill,
0 = one ray slot per (row, candidate) thread): device time of the whole build (CUDA events around vrad_build_transfers) and that
both forms leave bit-identical rows.  Writes gpurun_out/r02_k2_stream.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
res = {}
def run(name, scene, hier=None, reps=2):
    ref = None
    for flag in (1, 0):
        env = environment_from_scene(scene)
        if hier is not None:
            env.set_hierarchy(hier["parent"], hier["child1"], hier["child2"], hier["face"])
        env.set_option("k2_stream", flag)
        best = 1e30
        for _ in range(reps):
            torch.cuda.synchronize(); t0 = time.perf_counter(); nnz = env.build_transfers(scene.pvs); torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        rp, col, w = env.transfers_download()
        if ref is None: ref = (rp, col, w)
        same = bool(np.array_equal(rp, ref[0]) and np.array_equal(col, ref[1]) and np.array_equal(w, ref[2]))
        res[f"{name}_stream{flag}"] = {"wall_ms": best * 1e3, "device_ms": env.last_timing()[0], "nnz": int(nnz), "identical_to_stream1": same}
        print(name, "k2_stream", flag, "wall ms", best * 1e3, "nnz", nnz, "identical", same, flush=True)
        env.close()
run("C4_flat", scenes.multi_room())
os.environ["VRAD_K2_TOPDOWN"] = "0"
hs = scenes.multi_room_hier(nx=6, ny=5)
run("hier6x5_percandidate", hs, hs.meta["tree"], reps=1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r02_k2_stream.json", "w"), indent=1)
