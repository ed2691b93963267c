"""A/B timing of the flat K2 build (run from a tree root: `python tools/k2_ab.py`)."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
print("tree:", os.getcwd())
s = scenes.multi_room(); env = environment_from_scene(s)
for k in range(3):
    t = time.time(); nnz = env.build_transfers(s.pvs); dt = time.time() - t
    print(f"S2 flat build_transfers: nnz={nnz} wall {dt:.3f}s kernel_ms={env.last_timing()}")
env.close()
