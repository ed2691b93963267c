"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck): every kernel once, small sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
s = scenes.multi_room(nx=2, ny=1, boxes_per_room=8)
env = environment_from_scene(s)
a, b = scenes.shadow_segments(s, 20011)
env.test_lines(a, b); env.test_lines(a, b, sky_mode=1)
r = scenes.random_rays(s, 10007)
env.trace_rays(r["o"], r["d"], r["tmax"])
env.trace4_rays(r["o"][:, :4], r["d"][:, :4], np.zeros(4, np.float32), r["tmax"][:4])
sel = slice(0, None, 6)
env.patches_upload(s.patch_origin[sel], s.patch_normal[sel], s.patch_plane_dist[sel], s.patch_area[sel], s.patch_refl[sel], s.patch_cluster[sel], s.patch_flags[sel])
nnz = env.build_transfers(s.pvs)
n = s.patch_origin[sel].shape[0]
env.bounce(np.full((n, 3), 50.0, np.float32), 3)
env.bounce(np.full((n, 3), 50.0, np.float32), 50, early_out=True)
dirs = np.loadtxt(os.path.join(os.path.dirname(__file__), "..", "vrad_b200", "data", "anorms.txt"), dtype=np.float32)
env.set_sky_dirs(dirs)
L = np.zeros(3, scenes.LIGHT_DTYPE); L["start_fade"], L["end_fade"], L["cap_dist"] = 0.0, -1.0, 1e22
L[0]["type"] = 1; L[0]["origin"] = (200, 200, 400); L[0]["intensity"] = (1e6, 1e6, 1e6); L[0]["quadratic_attn"] = 1
L[1]["type"] = 3; L[1]["normal"] = (0, 0, -1); L[1]["intensity"] = (100, 100, 100)
L[2]["type"] = 5; L[2]["intensity"] = (10, 10, 10)
env.direct_light(s.patch_origin[::9], s.patch_normal[::9], L)
env.close()
print("sanitize target done, nnz", nnz)
# ---- widened rows: complete TestLineDoesHitSky, BSP point queries, K3 with trace flags, patch hierarchy ----
from vrad_b200.environment import Environment
sk = scenes.sky_room(n_boxes=8); m = sk.meta
g = Environment(); g.add_triangles(sk.tri_ids, sk.tri_verts, sk.tri_flags); g.set_triangle_colors(m["tri_colors"])
g.setup_acceleration_structure(); g.bsp_upload(m["bsp"]); g.process_sky_cameras(m["cams_origin"], m["cams_scale"])
sa, sb = scenes.sky_segments(sk, 5003)
for flags in (0, 1, 3, 7):
    g.test_lines_sky(sa, sb, flags, 7)
pts = np.ascontiguousarray(sa.T)
g.point_leafnum(pts); g.cluster_from_point(pts)
g.set_sky_dirs(dirs)
g.leafs_trace_to_sky(m["probe_mins"], m["probe_maxs"])
for flags in (1, 3):
    g.set_light_trace_flags(flags)
    g.direct_light(np.ascontiguousarray(pts[:999] * np.float32([1, 1, 0]) + np.float32([0, 0, 1])), np.tile(np.float32([0, 0, 1]), (999, 1)), L)
g.close()
hs = scenes.multi_room_hier(nx=2, ny=1, boxes_per_room=8); t = hs.meta["tree"]
h = environment_from_scene(hs)
h.set_hierarchy(t["parent"], t["child1"], t["child2"], t["face"])
hn = h.build_transfers(hs.pvs)
h.bounce(np.full((hs.n_patches, 3), 50.0, np.float32), 3)
h.close()
print("sanitize target (widened rows) done, hier nnz", hn)
