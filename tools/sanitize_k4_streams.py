"""compute-sanitizer target for the gather's stream builders and kernels (packed, block rows; single-GPU and work-item forms), small sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
s = scenes.multi_room(nx=2, ny=1, boxes_per_room=8)
env = environment_from_scene(s)
sel = slice(0, None, 3)
env.patches_upload(s.patch_origin[sel], s.patch_normal[sel], s.patch_plane_dist[sel], s.patch_area[sel], s.patch_refl[sel], s.patch_cluster[sel], s.patch_flags[sel])
nnz = env.build_transfers(s.pvs)
n = s.patch_origin[sel].shape[0]
e0 = np.full((n, 3), 50.0, np.float32)
for pack, items, rows in ((2, 0, 4), (3, 1, 4), (2, 0, 2), (3, 1, 2), (1, 0, 4), (1, 1, 4), (9, 0, 4), (0, 1, 4)):
    env.set_option("k4_items", items); env.set_option("k4_bk_rows", rows); env.set_option("k4_pack", pack)
    env.bounce(e0, 2); env.bounce(e0, 6)
    print(pack, items, rows, env.transfers_layout(), flush=True)
env.close()
print("sanitize k4 streams done, nnz", nnz)
