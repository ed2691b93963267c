import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
s = scenes.multi_room(); env = environment_from_scene(s); nnz = env.build_transfers(s.pvs)
rp, col, w = env.transfers_download()
lens = np.diff(rp)
print("rowlen percentiles", np.percentile(lens, [0, 1, 10, 50, 90, 99, 100]))
with open(sys.argv[1], "wb") as f:
    np.array([s.n_patches, nnz], np.int64).tofile(f); rp.tofile(f); col.tofile(f); w.tofile(f)
