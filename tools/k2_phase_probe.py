import os, sys, time
sys.path.insert(0, "/root/repo")
os.environ["VRAD_TIMING"] = "1"
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
s3 = scenes.outdoor()
for world, rank in ((8, 3), (1, 0)):
    env = environment_from_scene(s3, rank=rank, world=world)
    for _ in range(2):
        t0 = time.perf_counter(); nnz = env.build_transfers(s3.pvs); print("world", world, "wall", time.perf_counter() - t0, "nnz", nnz, flush=True)
    env.close()
